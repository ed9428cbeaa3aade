// Builds the benchmark plans and one kitchen-sink predicate with the C++ host mirror (rust-llkv_b200/host/llkv_gpu.hpp)
// and prints the flattened C-ABI arrays as hex, one line per array; tests/test_cpp_host.py compares them byte for byte
// with what the Python mirror flattens for the same trees.  No GPU, no library call.
#include <cstdio>

#include "../../rust-llkv_b200/host/llkv_gpu.hpp"

using namespace llkv;

template <typename T>
static void dump(const char* name, const std::vector<T>& v) {
  std::printf("%s %zu ", name, v.size());
  const unsigned char* b = reinterpret_cast<const unsigned char*>(v.data());
  for (size_t i = 0; i < v.size() * sizeof(T); ++i) std::printf("%02x", b[i]);
  std::printf("\n");
}
static void dump_program(const char* name, const Expr& e) {
  CompiledProgram p = ProgramCompiler(e).compile();
  std::printf("program %s\n", name);
  dump("ops", p.ops);
  dump("literals", p.literals);
  dump("nodes", p.pool.nodes);
  dump("list_roots", p.list_roots);
}
static void dump_aggs(const char* name, const std::vector<AggregateSpec>& specs) {
  FlatAggregates f = flatten_aggregates(specs);
  std::printf("aggregates %s\n", name);
  dump("specs", f.specs);
  dump("nodes", f.pool.nodes);
}

// field ids of the synthetic lineitem (rust-llkv_b200/llkv_b200/tpch.py)
enum : uint64_t { L_QUANTITY = 5, L_EXTENDEDPRICE = 6, L_DISCOUNT = 7, L_TAX = 8, L_RETURNFLAG = 9, L_LINESTATUS = 10, L_SHIPDATE = 11 };
static int32_t date32(int y, int m, int d) {  // days since 1970-01-01 (civil calendar)
  y -= m <= 2;
  const int era = (y >= 0 ? y : y - 399) / 400;
  const unsigned yoe = (unsigned)(y - era * 400);
  const unsigned doy = (153u * (unsigned)(m + (m > 2 ? -3 : 9)) + 2) / 5 + (unsigned)d - 1;
  const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + (int)doe - 719468;
}

int main() {
  const DataType dec152 = DataType::Decimal128(15, 2);
  // TPC-H Q6
  dump_program("q6", Expr::And({
      pred(L_SHIPDATE, Operator::Range(Bound::Included(Literal::Date32(date32(1994, 1, 1))), Bound::Excluded(Literal::Date32(date32(1995, 1, 1))))),
      pred(L_DISCOUNT, Operator::Range(Bound::Included(Literal::Decimal128(5, 2)), Bound::Included(Literal::Decimal128(7, 2)))),
      pred(L_QUANTITY, Operator::LessThan(Literal::Int128(24)))}));
  dump_aggs("q6", {{"revenue", AggregateKind::Sum(ScalarExpr::Column(L_EXTENDEDPRICE) * ScalarExpr::Column(L_DISCOUNT), dec152)}});
  // TPC-H Q1
  dump_program("q1", pred(L_SHIPDATE, Operator::LessThanOrEquals(Literal::Date32(date32(1998, 9, 2)))));
  {
    const ScalarExpr qty = ScalarExpr::Column(L_QUANTITY), price = ScalarExpr::Column(L_EXTENDEDPRICE), disc = ScalarExpr::Column(L_DISCOUNT),
                     tax = ScalarExpr::Column(L_TAX), one = ScalarExpr::Lit(1);
    const ScalarExpr disc_price = price * (one - disc), charge = disc_price * (one + tax);
    dump_aggs("q1", {{"sum_qty", AggregateKind::Sum(qty, dec152)},
                     {"sum_base_price", AggregateKind::Sum(price, dec152)},
                     {"sum_disc_price", AggregateKind::Sum(disc_price, DataType::Decimal128(38, 4))},
                     {"sum_charge", AggregateKind::Sum(charge, DataType::Decimal128(38, 6))},
                     {"avg_qty", AggregateKind::Avg(qty, dec152)},
                     {"avg_price", AggregateKind::Avg(price, dec152)},
                     {"avg_disc", AggregateKind::Avg(disc, DataType::Decimal128(38, 2))},
                     {"count_order", AggregateKind::CountStar()}});
  }
  // every Expr / Operator variant once: same-field AND (FusedAnd), OR, NOT, Compare, InList, IsNull, Literal, IN, strings
  dump_program("mixed", Expr::Not(Expr::Or({
      Expr::And({pred(1, Operator::GreaterThanOrEquals(-500)), pred(1, Operator::LessThanOrEquals(500))}),
      Expr::And({pred(2, Operator::In({1, 2, 3, -7})), pred(10, Operator::Equals("N")), pred(3, Operator::GreaterThan(-50.0)), pred(1, Operator::IsNotNull())}),
      Expr::Compare(ScalarExpr::Column(1) * ScalarExpr::Lit(2), CompareOp::LtEq, ScalarExpr::Column(4)),
      Expr::InList(ScalarExpr::Column(2), {ScalarExpr::Lit(5), ScalarExpr::Lit(Literal::Null()), ScalarExpr::Column(1)}, true),
      Expr::IsNull(ScalarExpr::Column(1) + ScalarExpr::Column(2)),
      Expr::Compare(ScalarExpr::Cast(ScalarExpr::Column(5), DataType::Float64()), CompareOp::Gt, ScalarExpr::Lit(Literal::Decimal128(1000000000, 4))),
      Expr::Literal(false)})));
  return 0;
}
