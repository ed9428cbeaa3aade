// GPU: TPC-H Q6 and a filtered COUNT/MIN/MAX driven entirely from the C++ host mirror (Context / Column / Program /
// Aggregation of rust-llkv_b200/host/llkv_gpu.hpp) over synthetic columns, checked against a scalar loop written here.
// Prints "ok ..." and exits 0, or the mismatch and exits 1.  Built and run by tests/test_cpp_host.py (-m gpu).
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../rust-llkv_b200/host/llkv_gpu.hpp"

using namespace llkv;
enum : uint64_t { L_QUANTITY = 5, L_EXTENDEDPRICE = 6, L_DISCOUNT = 7, L_SHIPDATE = 11 };

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 1000003;
  try {
    std::mt19937_64 rng(6);
    // Arrow Decimal128(15,2) values are 16-byte little-endian integers
    std::vector<__int128> qty(n), price(n), disc(n);
    std::vector<int32_t> ship(n);
    for (uint64_t i = 0; i < n; ++i) {
      const int64_t q = 1 + (int64_t)(rng() % 50);
      qty[i] = q * 100;
      price[i] = q * (90000 + (int64_t)(rng() % 120001));
      disc[i] = (int64_t)(rng() % 11);
      ship[i] = 8036 + (int32_t)(rng() % 2526);
    }
    Context ctx(0);
    const DataType dec = DataType::Decimal128(15, 2);
    Column c_qty(ctx, logical_field_id(1, L_QUANTITY), dec), c_price(ctx, logical_field_id(1, L_EXTENDEDPRICE), dec),
        c_disc(ctx, logical_field_id(1, L_DISCOUNT), dec), c_ship(ctx, logical_field_id(1, L_SHIPDATE), DataType::Date32());
    const uint64_t chunk = 4096;  // the reference's chunking for Decimal128 / Date32 columns (store/slicing.rs:155-166)
    for (uint64_t lo = 0; lo < n; lo += chunk) {
      const uint64_t m = std::min(chunk, n - lo);
      c_qty.append(&qty[lo], m, lo);
      c_price.append(&price[lo], m, lo);
      c_disc.append(&disc[lo], m, lo);
      c_ship.append(&ship[lo], m, lo);
    }
    c_qty.seal(); c_price.seal(); c_disc.seal(); c_ship.seal();

    const int32_t d0 = 8766, d1 = 9131;  // 1994-01-01, 1995-01-01
    Program q6(ctx, Expr::And({pred(L_SHIPDATE, Operator::Range(Bound::Included(Literal::Date32(d0)), Bound::Excluded(Literal::Date32(d1)))),
                               pred(L_DISCOUNT, Operator::Range(Bound::Included(Literal::Decimal128(5, 2)), Bound::Included(Literal::Decimal128(7, 2)))),
                               pred(L_QUANTITY, Operator::LessThan(Literal::Int128(24)))}));
    Aggregation agg(ctx, 1,
                    {{"revenue", AggregateKind::Sum(ScalarExpr::Column(L_EXTENDEDPRICE) * ScalarExpr::Column(L_DISCOUNT), dec)},
                     {"n", AggregateKind::CountStar()},
                     {"min_price", AggregateKind::Min(ScalarExpr::Column(L_EXTENDEDPRICE), dec)},
                     {"max_qty", AggregateKind::Max(ScalarExpr::Column(L_QUANTITY), dec)}});
    __int128 want_rev = 0, want_min = 0, want_max = 0;
    int64_t want_n = 0;
    for (uint64_t i = 0; i < n; ++i) {
      if (ship[i] >= d0 && ship[i] < d1 && disc[i] >= 5 && disc[i] <= 7 && qty[i] < 2400) {
        const __int128 prod = price[i] * disc[i];  // scale 4 -> back to scale 2, half away from zero (values >= 0)
        want_rev += (prod + 50) / 100;
        if (!want_n || price[i] < want_min) want_min = price[i];
        if (!want_n || qty[i] > want_max) want_max = qty[i];
        ++want_n;
      }
    }
    for (int rep = 0; rep < 3; ++rep) {  // the third run uses the kernel specialised on this plan shape
      agg.reset();
      agg.run(&q6, false, 0, n);
      const std::vector<GroupRow> rows = agg.finalize(1);
      if (rows.size() != 1) { std::printf("expected one row, got %zu\n", rows.size()); return 1; }
      const auto& v = rows[0].values;
      auto i128 = [](const llkv_agg_value& x) { return (__int128)(((unsigned __int128)x.hi << 64) | x.lo); };
      if (!v[0].valid || i128(v[0]) != want_rev || v[0].type != LLKV_PT_DECIMAL128) { std::printf("revenue mismatch (rep %d)\n", rep); return 1; }
      if ((int64_t)v[1].lo != want_n) { std::printf("count mismatch: %lld vs %lld\n", (long long)v[1].lo, (long long)want_n); return 1; }
      if (want_n && (i128(v[2]) != want_min || i128(v[3]) != want_max)) { std::printf("min/max mismatch\n"); return 1; }
    }
    const llkv_run_info info = agg.run_info();
    std::printf("ok rows=%llu selected=%lld revenue_raw=%lld lean=%u specialised=%u kernel_ms=%.4f\n", (unsigned long long)n, (long long)want_n,
                (long long)want_rev, info.used_fast_kernel, info.used_jit_kernel, info.last_kernel_ms);
    return 0;
  } catch (const Error& e) {
    std::printf("llkv error %d: %s\n", e.code, e.what());
    return 1;
  }
}
