#include "gdv_rope_temps.h"

namespace gdv {

bool IsRopeConsumerError(const Status& s) {
  if (s.code != GDV_NOT_IMPLEMENTED) return false;
  return s.msg.find("over it is not supported yet") != std::string::npos ||
         s.msg.find("(concat(...)) is not supported yet") != std::string::npos ||
         s.msg.find("over concat(...) is not supported yet") != std::string::npos;
}

namespace {

bool IsRopeFunction(const std::string& n) {
  return n == "concat" || n == "concatOperator" || n == "repeat" || n == "space" || n == "reverse" || n == "lpad" ||
         n == "rpad" || n == "replace";
}

// A node whose value is a rope: one of the functions above, or an if/else that may yield one.
bool IsRope(const Node& node) {
  if (!node.return_type().is_varlen()) return false;
  if (node.kind() == NodeKind::kFunction) return IsRopeFunction(static_cast<const FunctionNode&>(node).name());
  if (node.kind() == NodeKind::kIf) {
    const auto& n = static_cast<const IfNode&>(node);
    return IsRope(*n.then_node()) || IsRope(*n.else_node());
  }
  return false;
}

struct Extractor {
  RopeTemps* out;
  bool in_temp = false;  // walking the expression of a temp

  // `rope_ok`: the parent can read a rope (it is the output root, a concat, or an if/else in such a position)
  NodePtr Walk(const NodePtr& node, bool rope_ok) {
    if (IsRope(*node) && !rope_ok) {
      // a consumer inside the arguments of a rope that is itself being materialised stays where it is: the
      // internal Projector that evaluates the temp is built the same way and takes care of it (next level)
      if (in_temp) return node;
      in_temp = true;
      const NodePtr inner = Walk(node, true);
      in_temp = false;
      const std::string text = inner->ToString();
      for (size_t k = 0; k < out->temps.size(); ++k)
        if (out->temps[k]->ToString() == text && out->fields[k].type == node->return_type())
          return std::make_shared<FieldNode>(out->fields[k].name, out->fields[k].type);
      Field f{"__gdv_rope_" + std::to_string(out->temps.size()), node->return_type()};
      out->temps.push_back(std::make_shared<Expression>(inner, f));
      out->fields.push_back(f);
      return std::make_shared<FieldNode>(f.name, f.type);
    }
    switch (node->kind()) {
      case NodeKind::kFunction: {
        const auto& fn = static_cast<const FunctionNode&>(*node);
        const bool reads_ropes = IsRopeFunction(fn.name());  // its string arguments may be ropes themselves
        NodeVector kids;
        bool changed = false;
        for (const auto& c : fn.children()) {
          kids.push_back(Walk(c, reads_ropes && (fn.name() == "concat" || fn.name() == "concatOperator")));
          changed = changed || kids.back() != c;
        }
        return changed ? std::make_shared<FunctionNode>(fn.name(), std::move(kids), fn.return_type()) : node;
      }
      case NodeKind::kIf: {
        const auto& n = static_cast<const IfNode&>(*node);
        const NodePtr c = Walk(n.condition(), false), t = Walk(n.then_node(), rope_ok), e = Walk(n.else_node(), rope_ok);
        if (c == n.condition() && t == n.then_node() && e == n.else_node()) return node;
        return std::make_shared<IfNode>(c, t, e, n.return_type());
      }
      case NodeKind::kBoolean: {
        const auto& n = static_cast<const BooleanNode&>(*node);
        NodeVector kids;
        bool changed = false;
        for (const auto& c : n.children()) {
          kids.push_back(Walk(c, false));
          changed = changed || kids.back() != c;
        }
        return changed ? std::make_shared<BooleanNode>(n.op(), std::move(kids)) : node;
      }
      case NodeKind::kIn: {
        const auto& n = static_cast<const InNode&>(*node);
        const NodePtr c = Walk(n.child(), false);
        return c == n.child() ? node : std::make_shared<InNode>(c, n.value_type(), n.ints(), n.strs());
      }
      default:
        return node;
    }
  }
};

}  // namespace

bool ExtractRopes(const NodePtr& root, bool root_is_output, RopeTemps* out, NodePtr* rewritten) {
  Extractor x{out};
  *rewritten = x.Walk(root, root_is_output);
  return true;
}

}  // namespace gdv
