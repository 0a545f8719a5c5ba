#include "gdv_staging.h"

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "gdv_runtime.h"

namespace gdv {

namespace {

constexpr size_t kSlotBytes = size_t(8) << 20;   // one DMA piece: ~150 us at PCIe Gen5 x16
constexpr int kSlots = 6;                        // pieces in flight (memcpy of i+1.. while DMA of i)
constexpr size_t kDirectBelow = size_t(1) << 20; // smaller copies are not worth the hand-offs

std::atomic<long long> g_staged_bytes{0};

// A few threads that copy slices of one piece in parallel (one memcpy stream does ~10 GB/s, the
// link wants ~55).  Started on first use, parked on a condition variable in between.
class CopyPool {
 public:
  static CopyPool& Get() {
    static CopyPool* p = new CopyPool();  // never destroyed: threads may outlive static destructors
    return *p;
  }
  void Copy(void* dst, const void* src, size_t bytes) {
    const int n = static_cast<int>(workers_.size()) + 1;
    if (bytes < (size_t(256) << 10) || n == 1) {
      std::memcpy(dst, src, bytes);
      return;
    }
    const size_t slice = ((bytes + n - 1) / n + 4095) & ~size_t(4095);
    {
      std::lock_guard<std::mutex> g(mu_);
      dst_ = static_cast<char*>(dst);
      src_ = static_cast<const char*>(src);
      bytes_ = bytes;
      slice_ = slice;
      pending_ = n - 1;
      pending_pub_.store(n - 1, std::memory_order_release);
      ++epoch_;
      epoch_pub_.store(epoch_, std::memory_order_release);
    }
    cv_.notify_all();
    Slice(0);  // the calling thread takes the first slice
    for (int spin = 0; spin < 20000; ++spin)  // the others finish within microseconds of this one
      if (pending_pub_.load(std::memory_order_acquire) == 0) break;
    std::unique_lock<std::mutex> g(mu_);
    done_.wait(g, [this] { return pending_ == 0; });
  }

 private:
  CopyPool() {
    int want = 8;
    if (const char* e = std::getenv("GDV_STAGE_THREADS")) want = std::atoi(e);
    const int hw = static_cast<int>(std::thread::hardware_concurrency());
    if (hw > 0 && want > hw) want = hw;
    for (int t = 1; t < want; ++t) workers_.emplace_back([this, t] { Loop(t); });
    for (auto& w : workers_) w.detach();
  }
  void Slice(int t) {
    const size_t b = slice_ * static_cast<size_t>(t);
    if (b < bytes_) std::memcpy(dst_ + b, src_ + b, std::min(slice_, bytes_ - b));
  }
  void Loop(int t) {
    unsigned long long seen = 0;
    for (;;) {
      // A host batch is several copies back to back (one per column, then the results): after a job the
      // worker polls for the next one for ~100 us before it goes to sleep, because waking a sleeping thread
      // costs tens of microseconds — as much as copying its whole slice.
      bool got = false;
      for (int spin = 0; spin < 20000 && !got; ++spin) {
        if (epoch_pub_.load(std::memory_order_acquire) != seen) got = true;
        else if ((spin & 63) == 63) std::this_thread::yield();
      }
      {
        std::unique_lock<std::mutex> g(mu_);
        if (!got) cv_.wait(g, [&] { return epoch_ != seen; });
        seen = epoch_;
      }
      Slice(t);
      std::lock_guard<std::mutex> g(mu_);
      pending_pub_.store(pending_ - 1, std::memory_order_release);
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0, slice_ = 0;
  int pending_ = 0;
  unsigned long long epoch_ = 0;
  std::atomic<unsigned long long> epoch_pub_{0};  // epoch_, readable without the lock by polling workers
  std::atomic<int> pending_pub_{0};               // pending_, for the caller's short poll
};

// The pinned slots of one device.  One transfer at a time uses the ring (the link is shared anyway).
struct Ring {
  std::mutex mu;
  bool ready = false;
  void* slot[kSlots] = {nullptr};
  CUevent ev[kSlots] = {nullptr};
  bool busy[kSlots] = {false};
  int next = 0;

  Status Init() {
    if (ready) return Status::OK();
    const DriverApi& d = Driver();
    for (int i = 0; i < kSlots; ++i) {
      Status s = CuCheck(d.MemHostAlloc(&slot[i], kSlotBytes, 0), "cuMemHostAlloc(staging slot)");
      if (!s.ok()) return s;
      s = CuCheck(d.EventCreate(&ev[i], CU_EVENT_DISABLE_TIMING), "cuEventCreate");
      if (!s.ok()) return s;
    }
    ready = true;
    return Status::OK();
  }
};

std::mutex g_rings_mu;
Ring* g_rings[64] = {nullptr};

Ring* RingFor(Device* dev) {
  std::lock_guard<std::mutex> g(g_rings_mu);
  Ring*& r = g_rings[dev->ordinal()];
  if (r == nullptr) r = new Ring();
  return r;
}

}  // namespace

long long StagedBytes() { return g_staged_bytes.load(); }

bool IsPageableHost(const void* p) {
  const DriverApi& d = Driver();
  if (!d.loaded) return false;
  unsigned int type = 0;
  const CUresult r = d.PointerGetAttribute(&type, CU_POINTER_ATTRIBUTE_MEMORY_TYPE,
                                           reinterpret_cast<CUdeviceptr>(p));
  // memory the driver has never seen: cuPointerGetAttribute fails with INVALID_VALUE
  return r != CUDA_SUCCESS;
}

Status StagedHtoD(Device* dev, CUdeviceptr dst, const void* src, size_t bytes, CUstream stream) {
  const DriverApi& d = Driver();
  if (bytes == 0) return Status::OK();
  if (bytes < kDirectBelow || !IsPageableHost(src))
    return CuCheck(d.MemcpyHtoDAsync(dst, src, bytes, stream), "cuMemcpyHtoDAsync");
  Ring* ring = RingFor(dev);
  std::lock_guard<std::mutex> lock(ring->mu);
  Status st = ring->Init();
  if (!st.ok()) return st;
  CopyPool& pool = CopyPool::Get();
  const char* s = static_cast<const char*>(src);
  for (size_t off = 0; off < bytes; off += kSlotBytes) {
    const size_t n = std::min(kSlotBytes, bytes - off);
    const int i = ring->next;
    ring->next = (i + 1) % kSlots;
    if (ring->busy[i]) {  // the DMA that last read this slot must be done
      st = CuCheck(d.EventSynchronize(ring->ev[i]), "cuEventSynchronize(staging slot)");
      if (!st.ok()) return st;
      ring->busy[i] = false;
    }
    pool.Copy(ring->slot[i], s + off, n);
    st = CuCheck(d.MemcpyHtoDAsync(dst + off, ring->slot[i], n, stream), "cuMemcpyHtoDAsync(staged)");
    if (!st.ok()) return st;
    st = CuCheck(d.EventRecord(ring->ev[i], stream), "cuEventRecord");
    if (!st.ok()) return st;
    ring->busy[i] = true;
  }
  g_staged_bytes.fetch_add(static_cast<long long>(bytes));
  return Status::OK();
}

Status StagedDtoH(Device* dev, void* dst, CUdeviceptr src, size_t bytes, CUstream stream) {
  const DriverApi& d = Driver();
  if (bytes == 0) return Status::OK();
  if (bytes < kDirectBelow || !IsPageableHost(dst))
    return CuCheck(d.MemcpyDtoHAsync(dst, src, bytes, stream), "cuMemcpyDtoHAsync");
  Ring* ring = RingFor(dev);
  std::lock_guard<std::mutex> lock(ring->mu);
  Status st = ring->Init();
  if (!st.ok()) return st;
  CopyPool& pool = CopyPool::Get();
  char* out = static_cast<char*>(dst);
  // every slot may still be read by an earlier H2D: wait for those first
  for (int i = 0; i < kSlots; ++i)
    if (ring->busy[i]) {
      st = CuCheck(d.EventSynchronize(ring->ev[i]), "cuEventSynchronize(staging slot)");
      if (!st.ok()) return st;
      ring->busy[i] = false;
    }
  const size_t pieces = (bytes + kSlotBytes - 1) / kSlotBytes;
  // DMA of piece p + depth is enqueued before piece p is copied out of its slot
  const size_t depth = static_cast<size_t>(kSlots) - 1;
  auto issue = [&](size_t p) -> Status {
    const size_t off = p * kSlotBytes, n = std::min(kSlotBytes, bytes - off);
    const int i = static_cast<int>(p % kSlots);
    Status s2 = CuCheck(d.MemcpyDtoHAsync(ring->slot[i], src + off, n, stream), "cuMemcpyDtoHAsync(staged)");
    if (!s2.ok()) return s2;
    return CuCheck(d.EventRecord(ring->ev[i], stream), "cuEventRecord");
  };
  for (size_t p = 0; p < std::min(depth, pieces); ++p) {
    st = issue(p);
    if (!st.ok()) return st;
  }
  for (size_t p = 0; p < pieces; ++p) {
    const size_t off = p * kSlotBytes, n = std::min(kSlotBytes, bytes - off);
    const int i = static_cast<int>(p % kSlots);
    st = CuCheck(d.EventSynchronize(ring->ev[i]), "cuEventSynchronize(staging slot)");
    if (!st.ok()) return st;
    pool.Copy(out + off, ring->slot[i], n);
    if (p + depth < pieces) {
      st = issue(p + depth);
      if (!st.ok()) return st;
    }
  }
  ring->next = 0;
  g_staged_bytes.fetch_add(static_cast<long long>(bytes));
  return Status::OK();
}

}  // namespace gdv
