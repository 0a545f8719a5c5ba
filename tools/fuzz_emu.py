"""Differential fuzzing of the fuser against the oracle on a CPU-only box (functional simulator, tests/emu).
  python tools/fuzz_emu.py FIRST LAST [proj|filt|str]     seeds FIRST..LAST-1, failures printed with their seed
Reuses the tree generator of tests/test_random_trees.py (other seeds, other sizes) and adds a string-filter
fuzzer: LIKE / substr / upper / IN conjunctions over comment-like text with non-ASCII rows, dense and rare
keys — the shapes the key-driven filter kernel is planned from.  Test infrastructure only."""
import os
import subprocess
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(first, last, what):
    import numpy as np
    import pyarrow as pa
    import cases
    import gandiva_b200 as gandiva
    import oracle
    import test_random_trees as T
    from helpers import assert_arrays_match
    bad = skipped = 0
    if what in ("rope", "ropefilt"):
        # ropes read by arbitrary consumers (materialised through a temporary column, csrc/gdv_rope_temps.h);
        # trees that nest a consumer inside another rope's arguments are refused at Make() and counted as skipped
        T.SKIP -= {"concat", "concatOperator", "reverse"}
        what_kind = "proj" if what == "rope" else "filt"
    else:
        what_kind = what
    for seed in range(first, last):
        if os.environ.get("FUZZ_VERBOSE"):
            print("seed %d" % seed, flush=True)
        try:
            rng = np.random.default_rng(77_000 + seed)
            b = gandiva.TreeExprBuilder()
            g = T.TreeGen(gandiva, b, rng)
            if what_kind == "proj":
                out_types = [T.TYPES[int(rng.integers(len(T.TYPES)))] for _ in range(int(rng.integers(1, 5)))]
                roots = [g.gen(t, int(rng.integers(2, 6))) for t in out_types]
                exprs = [b.make_expression(r, pa.field("o%d" % i, t)) for i, (r, t) in enumerate(zip(roots, out_types))]
                p = gandiva.make_projector(T.SCHEMA, exprs, None)
                for n in (int(rng.integers(1, 70)), int(rng.integers(200, 700)) if what == "rope" else int(rng.integers(1000, 6000))):
                    batch = cases.random_batch(T.SCHEMA, n, seed=seed + n, null_prob=float(rng.choice([0.0, 0.12, 0.5])),
                                               offset=int(rng.integers(0, 9)))
                    got = p.evaluate(batch)
                    want = oracle.project(roots, out_types, batch, threads=2)
                    for i, (gv, wv) in enumerate(zip(got, want)):
                        assert_arrays_match(gv, wv, "seed %d n=%d out %d: %s" % (seed, n, i, roots[i]))
            elif what_kind == "filt":
                cond = g.gen(T.B, int(rng.integers(2, 4 if what == "ropefilt" else 6)))   # (deep rope trees: many plan levels)
                cfg = gandiva.Configuration(string_scan=4) if seed % 3 == 0 else None
                f = gandiva.make_filter(T.SCHEMA, b.make_condition(cond), cfg)
                for n in (int(rng.integers(1, 70)), int(rng.integers(200, 700)) if what == "ropefilt" else int(rng.integers(2000, 12000))):
                    batch = cases.random_batch(T.SCHEMA, n, seed=seed + n, null_prob=float(rng.choice([0.0, 0.12, 0.5])),
                                               offset=int(rng.integers(0, 9)))
                    got = f.evaluate(batch).to_array().to_numpy().astype(np.uint64)
                    want = oracle.filter_indices(cond, batch, threads=2)
                    assert np.array_equal(got, want), "seed %d n=%d: %s" % (seed, n, cond)
            else:
                S, B = pa.string(), pa.bool_()
                words = ["special", "requests", "packages", "deposits", "fur", "the", "quick", "日本", "ß", "é", "x", "_", "%",
                         "Special", "REQUESTS", " ", "ab", "aab", "aaab", "cial re", "unusual", "express"]
                nw = int(rng.integers(2, 9))
                n = int(rng.integers(1, 20000))
                rows = []
                pnull = float(rng.choice([0.0, 0.05, 0.4]))
                dense = rng.random() < 0.3
                for _ in range(n):
                    if rng.random() < pnull:
                        rows.append(None)
                        continue
                    k = int(rng.integers(0, nw + 1))
                    pool = words[:6] if dense else words
                    rows.append("".join(str(rng.choice(pool)) + (" " if rng.random() < 0.7 else "") for _ in range(k)))
                schema = pa.schema([("c", S), ("k", pa.int32())])
                batch = pa.record_batch([pa.array(rows, S), pa.array(rng.integers(0, 100, n, dtype=np.int32),
                                                                      mask=rng.random(n) < pnull)], schema=schema)
                off = int(rng.integers(0, 7))
                if off and n > off:
                    batch = batch.slice(off)
                fc, fk = b.make_field(schema.field(0)), b.make_field(schema.field(1))
                fn = b.make_function

                def pat():
                    parts = [str(rng.choice(words)) for _ in range(int(rng.integers(1, 4)))]
                    style = int(rng.integers(0, 6))
                    if style == 0:
                        return "%" + "%".join(parts) + "%"
                    if style == 1:
                        return parts[0] + "%"
                    if style == 2:
                        return "%" + parts[0]
                    if style == 3:
                        return "%" + parts[0] + "_" + (parts[1] if len(parts) > 1 else "") + "%"
                    if style == 4:
                        return "".join(parts)
                    return "%" + parts[0][: max(1, len(parts[0]) // 2)] + "%"

                def leaf():
                    r = rng.random()
                    if r < 0.45:
                        return fn("like", [fc, b.make_literal(pat(), S)], B)
                    if r < 0.55:
                        return fn("like", [fn("upper", [fc], S), b.make_literal(pat().upper(), S)], B)
                    if r < 0.65:
                        return fn("like", [fn("substr", [fc, b.make_literal(int(rng.integers(1, 5)), pa.int64()),
                                                         b.make_literal(int(rng.integers(0, 30)), pa.int64())], S),
                                           b.make_literal(pat(), S)], B)
                    if r < 0.75:
                        return fn("less_than", [fk, b.make_literal(int(rng.integers(0, 100)), pa.int32())], B)
                    if r < 0.85:
                        return b.make_in_expression(fc, [str(rng.choice(words)) for _ in range(3)], S)
                    if r < 0.92:
                        return fn("starts_with", [fc, b.make_literal(str(rng.choice(words)), S)], B)
                    return fn("not", [fn("like", [fc, b.make_literal(pat(), S)], B)], B)
                kids = [leaf() for _ in range(int(rng.integers(1, 4)))]
                cond = kids[0] if len(kids) == 1 else (b.make_and(kids) if rng.random() < 0.7 else b.make_or(kids))
                cfg = gandiva.Configuration(string_scan=4) if seed % 4 == 0 else None
                f = gandiva.make_filter(schema, b.make_condition(cond), cfg)
                got = f.evaluate(batch).to_array().to_numpy().astype(np.uint64)
                want = oracle.filter_indices(cond, batch, threads=2)
                assert np.array_equal(got, want), "seed %d n=%d key_driven=%s: %s" % (seed, n, f.kernel_info.get("key_driven"), cond)
        except pa.ArrowNotImplementedError as e:
            skipped += 1
            if os.environ.get("FUZZ_VERBOSE"):
                print("refused %s seed %d: %s" % (what, seed, str(e)[:300]), flush=True)
        except Exception as e:  # noqa: BLE001 - report and go on
            bad += 1
            msg = traceback.format_exc().strip().splitlines()
            print("FAIL %s seed %d: %s" % (what, seed, " | ".join(msg[-3:])[:1500]), flush=True)
    print("done %s %d..%d: %d failures, %d refused at Make()" % (what, first, last, bad, skipped), flush=True)


def main():
    first, last = int(sys.argv[1]), int(sys.argv[2])
    what = sys.argv[3] if len(sys.argv) > 3 else "proj"
    if os.environ.get("GDV_EMU") == "1":
        child(first, last, what)
        return
    import emu
    sys.exit(subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=emu.env()).returncode)


if __name__ == "__main__":
    main()
