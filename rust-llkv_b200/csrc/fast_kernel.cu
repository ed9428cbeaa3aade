// fast_kernel.cu — the lean scan -> filter -> MVCC -> project -> aggregate kernel for sm_100a.
//
// Runs the FastOp program the compiler lowers (compiler.cpp: lower_fast) when a plan has no NULLs and range analysis over
// the columns' min/max statistics proves that every value fits 64 bits.  Everything else runs on the general
// interpreter (scan_kernel.cu); both produce identical accumulator states.
//
// Shape of one CTA: NC consumer threads + one producer warp.
//   * The producer's elected lane keeps `stages` row tiles in flight with cp.async.bulk (TMA 1-D bulk copies, one per
//     column per tile) and full/empty mbarriers per stage; consumers never wait on HBM and there is no __syncthreads
//     in the tile loop.
//   * Each consumer thread owns R rows of the tile (row = r * NC + thread, so a warp's shared-memory reads of 4/8/16 B
//     elements are conflict free).  Typed predicate leaves compare straight out of the tile into a bit mask; the
//     projection arithmetic is an accumulator machine: one i64 per row in registers, the other operand read from a
//     column tile, a literal or a tile-sized temporary in shared memory.
//   * Aggregates are group-major: for each CTA-local group slot present in the warp, a lane folds its own R rows, the
//     warp reduces once (redux.sync on 24-bit limbs for integer sums, butterfly shuffles for f64 / min / max) and lane 0
//     adds the result to the warp's private accumulator row in shared memory: no shared-memory atomics, no per-thread
//     accumulators.
//   * At the end the warps' rows are combined and folded into the global group table with one atomic per word.
#include "device_util.cuh"
#include "plan.h"

namespace llkv {

#define FULL 0xffffffffu

__device__ __forceinline__ double f_as_f64(i64 v) { return __longlong_as_double(v); }
__device__ __forceinline__ i64 f_bits(double d) { return __double_as_longlong(d); }

// rows without a CTA-local group slot (more groups than slots) and values too large for the per-warp i64 partials go
// straight to the global table, one atomic per row: rare, kept out of line
static __device__ __noinline__ void slow_accumulate(const Plan& p, uint32_t op, uint32_t flags, u64 K, i64 v, u64 row, uint32_t gword,
                                                    uint32_t& errbits) {
  u64* w = &p.gwords[global_slot(p, K, false, errbits) * p.n_gwords + gword];
  switch (op) {
    case FO_COUNT_STAR: case FO_COUNT: atomicAdd(w, 1ull); break;
    case FO_FIRSTROW: case FO_FIRSTVALID: case FO_FIRSTNAN: atomicMin(w, row); break;
    case FO_SUM:
      if (flags & 0x80) gadd_sum_i128(w, (i128)v);
      else gadd_sum_i64(w, (i128)v);
      break;
    case FO_FSUM: atomicAdd(reinterpret_cast<double*>(w), f_as_f64(v)); break;
    case FO_MIN_I: atomicMin(w, enc_i64(v)); break;
    case FO_MAX_I: atomicMax(w, enc_i64(v)); break;
    case FO_MIN_F: atomicMin(w, enc_f64(f_as_f64(v))); break;
    default: atomicMax(w, enc_f64(f_as_f64(v))); break;
  }
}

__device__ __forceinline__ void mbar_wait_backoff(void* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
  }
}

template <int R>
__global__ void __launch_bounds__(R >= 8 ? 160 : 288, R >= 8 ? 3 : 2) fast_scan_kernel(const Plan* __restrict__ gplan) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int NC = blockDim.x - 32;  // consumer threads
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int n_cwarps = NC >> 5;
  const bool is_producer = warp == n_cwarps;

  Plan& p = *reinterpret_cast<Plan*>(smem);
  {
    const uint32_t n4 = sizeof(Plan) / 4;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(gplan);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
    for (uint32_t i = tid; i < n4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();

  u64* const full_bar = reinterpret_cast<u64*>(smem + p.smem_bar_off);
  u64* const empty_bar = full_bar + p.stages;
  unsigned char* const stage0 = smem + p.smem_stage_off;
  u64* const wacc = reinterpret_cast<u64*>(smem + p.smem_acc_off);  // [consumer warp][group slot][word]
  u64* const tbl = reinterpret_cast<u64*>(smem + p.smem_tbl_off);   // CTA-local group keys

  const uint32_t FG = p.fast_groups;
  const uint32_t NFW = p.n_fast_words;
  const uint32_t S = p.stages;
  for (uint32_t i = tid; i < (uint32_t)n_cwarps * FG * NFW; i += blockDim.x) {
    const uint32_t w = i % NFW;
    const uint8_t k = p.fast[w].kind;
    u64 init = 0;
    if (k == FK_MIN) init = ~0ull;
    wacc[i] = init;
  }
  for (uint32_t g = tid; g < FG; g += blockDim.x) tbl[g] = (p.n_keys == 0) ? 0ull : kEmptyKey;
  if (tid == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], (uint32_t)n_cwarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const u64 n_tiles = p.n_tiles;
  const uint32_t T = p.tile_rows;
  const u64 my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  uint32_t errbits = 0;

  if (is_producer) {
    // ---------------------------------------------------------------- producer: TMA bulk copies, `S` tiles in flight
    if (lane == 0) {
      for (u64 li = 0; li < my_tiles; ++li) {
        const uint32_t s = (uint32_t)(li % S);
        if (li >= S) mbar_wait_backoff(&empty_bar[s], (uint32_t)(((li / S) - 1) & 1));
        const u64 tile = p.first_tile + blockIdx.x + li * gridDim.x;
        unsigned char* sb = stage0 + (size_t)s * p.stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], p.tx_bytes);
        for (uint32_t c = 0; c < p.n_cols; ++c) {
          const ColDesc& cd = p.cols[c];
          const uint32_t bytes = T * cd.elem_bytes;
          bulk_g2s(sb + cd.smem_off, reinterpret_cast<const unsigned char*>(cd.base) + tile * (u64)bytes, bytes, &full_bar[s]);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- consumers
    u64* const my_acc = wacc + (size_t)warp * FG * NFW;
    i64* const tmp_base = reinterpret_cast<i64*>(smem + p.smem_tmp_off);
    for (u64 li = 0; li < my_tiles; ++li) {
      const uint32_t s = (uint32_t)(li % S);
      const u64 tile = p.first_tile + blockIdx.x + li * gridDim.x;
      const u64 row0 = tile * (u64)T;
      mbar_wait_backoff(&full_bar[s], (uint32_t)((li / S) & 1));
      const unsigned char* sb = stage0 + (size_t)s * p.stage_bytes;

      // per-row state kept in registers across ops: the accumulator, the active bit and the CTA-local group slot
      i64 acc[R];
      int slot[R];
      unsigned actm = 0;       // bit r: row r of this thread is selected
      unsigned negm = 0;       // bit r: selected row without a CTA-local group slot (goes to the global table directly)
      unsigned present = 1u;   // CTA-local group slots present among this warp's selected rows (ungrouped: slot 0)
      bool has_slow = false;   // warp-uniform: some lane has a row in negm
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const u64 row = row0 + (u64)r * NC + tid;
        if (row >= p.row_begin && row < p.row_end) actm |= 1u << r;
        acc[r] = 0;
        slot[r] = 0;
      }

      // one column's R values as sign/zero-extended i64.  The lean kernel knows four physical layouts.
      auto load_col = [&](uint32_t col, uint32_t kind, i64 (&out)[R]) {
        const unsigned char* base = sb + p.cols[col].smem_off;
        if (kind == LKF_8) {
#pragma unroll
          for (int r = 0; r < R; ++r) out[r] = reinterpret_cast<const i64*>(base)[r * NC + tid];
        } else if (kind == LKF_4) {
#pragma unroll
          for (int r = 0; r < R; ++r) out[r] = reinterpret_cast<const int*>(base)[r * NC + tid];
        } else if (kind == LKF_16) {  // Decimal128 proven to hold sign-extended i64 values: the low half is the value
#pragma unroll
          for (int r = 0; r < R; ++r) out[r] = reinterpret_cast<const i64*>(base)[2 * (r * NC + tid)];
        } else if (kind == LKF_1) {
#pragma unroll
          for (int r = 0; r < R; ++r) out[r] = base[r * NC + tid];
        } else {  // LKF_S1: one-byte strings as packed keys
#pragma unroll
          for (int r = 0; r < R; ++r) out[r] = (i64)(((u64)base[r * NC + tid] << 56) | 1ull);
        }
      };
      // acc = acc op other  (rev: other op acc)
      auto binop = [&](uint32_t opr, const i64 (&o)[R]) {
        const bool rev = (opr & FB_REV) != 0;
        switch (opr & 0x7f) {
          case FB_ADD:
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] += o[r];
            break;
          case FB_SUB:
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = rev ? o[r] - acc[r] : acc[r] - o[r];
            break;
          case FB_MUL:
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] *= o[r];
            break;
          case FB_MUL32:
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = (i64)(int)acc[r] * (i64)(int)o[r];
            break;
          case FB_ADD_CK: case FB_SUB_CK: case FB_MUL_CK:
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const i64 a = rev ? o[r] : acc[r], b = rev ? acc[r] : o[r];
              i64 c;
              const uint32_t k = opr & 0x7f;
              const bool ok = k == FB_ADD_CK ? add_ck(a, b, c) : k == FB_SUB_CK ? sub_ck(a, b, c) : mul_ck(a, b, c);
              if (!ok && ((actm >> r) & 1u)) errbits |= FLAG_NARROW_FAIL;
              acc[r] = c;
            }
            break;
          case FB_ADD_F:
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = f_bits(f_as_f64(acc[r]) + f_as_f64(o[r]));
            break;
          case FB_SUB_F:
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = f_bits(rev ? f_as_f64(o[r]) - f_as_f64(acc[r]) : f_as_f64(acc[r]) - f_as_f64(o[r]));
            break;
          default:
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = f_bits(f_as_f64(acc[r]) * f_as_f64(o[r]));
            break;
        }
      };
      // packed GROUP BY key of row r
      auto row_key = [&](int r) -> u64 {
        u64 K = 0;
        int shift = 0;
        for (uint32_t k = 0; k < p.n_keys; ++k) {
          const unsigned char* base = sb + p.cols[p.key_col[k]].smem_off;
          const uint32_t i = (uint32_t)(r * NC + tid);
          const uint32_t kind = p.key_load[k];
          i64 v;
          if (kind == LKF_S1) v = (i64)(((u64)base[i] << 56) | 1ull);
          else if (kind == LKF_1) v = base[i];
          else if (kind == LKF_4) v = reinterpret_cast<const int*>(base)[i];
          else v = reinterpret_cast<const i64*>(base)[i];
          const int bits = p.key_bits[k];
          u64 f;
          if (p.key_kind[k] == KK_STR) {
            const int L = p.key_strlen[k];
            f = (L ? (((u64)v >> (64 - 8 * L)) << 3) : 0ull) | ((u64)v & 7ull);
          } else {
            f = (u64)v - p.key_min[k];
          }
          if (p.single_wide_key) K = (u64)v;
          else K |= (bits == 64 ? f : (f & ((1ull << bits) - 1))) << shift;
          shift += bits;
        }
        return K;
      };
      auto slow_rows = [&](uint32_t op, uint32_t flags, unsigned rows, uint32_t gword) {  // warp-uniform guard at the call site
#pragma unroll
        for (int r = 0; r < R; ++r)
          if ((rows >> r) & 1u) slow_accumulate(p, op, flags, row_key(r), acc[r], row0 + (u64)r * NC + tid, gword, errbits);
      };

      uint32_t pc = 0;
      FInstr in = p.fcode[0];
      while (true) {
        // optional operand pre-load fused into the instruction: acc = literal / column / temporary
        if (in.d) {
          if (in.d == 2) load_col(in.e, in.f, acc);
          else if (in.d == 1) {
            const i64 v = (i64)p.lits[in.e].lo;
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = v;
          } else {
            const i64* t = tmp_base + (size_t)in.e * T;
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = t[r * NC + tid];
          }
        }
        bool done = false;
        switch (in.op) {
          case FO_LEAF: {
            // consecutive leaves run back to back without going through the dispatcher again
            do {
              const Lit lo = p.lits[in.c], hi = p.lits[in.c + 1];
              const unsigned char* base = sb + p.cols[in.a].smem_off;
              unsigned m = 0;
              if (in.b == LKF_4) {
                const int l = (int)(i64)lo.lo, h = (int)(i64)hi.lo;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  const int v = reinterpret_cast<const int*>(base)[r * NC + tid];
                  m |= (unsigned)(v >= l && v <= h) << r;
                }
              } else if ((in.b == LKF_8 || in.b == LKF_16) && !in.g) {
                const i64 l = (i64)lo.lo, h = (i64)hi.lo;
                const int stride = in.b == LKF_16 ? 2 : 1;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  const i64 v = reinterpret_cast<const i64*>(base)[stride * (r * NC + tid)];
                  m |= (unsigned)(v >= l && v <= h) << r;
                }
              } else {  // unsigned 8-byte / 1-byte kinds
                i64 v[R];
                load_col(in.a, in.b, v);
#pragma unroll
                for (int r = 0; r < R; ++r) m |= (unsigned)((u64)v[r] >= lo.lo && (u64)v[r] <= hi.lo) << r;
              }
              actm &= m;
              in = p.fcode[++pc];
            } while (in.op == FO_LEAF);
            continue;
          }
          case FO_MVCC: {
            // RowVersion::is_visible_for (llkv-transaction/src/mvcc.rs:282-334)
            const u64* cbase = reinterpret_cast<const u64*>(sb + p.cols[in.a].smem_off);
            const u64* dbase = reinterpret_cast<const u64*>(sb + p.cols[in.b].smem_off);
            unsigned m = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const u64 cb = cbase[r * NC + tid], db = dbase[r * NC + tid];
              bool vis;
              if (cb == 1ull && db == ~0ull) vis = p.snapshot_id >= 1ull;  // auto-committed, never deleted: rules 2-4
              else {
                bool c_committed = cb != ~0ull, d_committed = db != ~0ull;
                for (uint32_t k = 0; k < p.n_noncommitted; ++k) {
                  const u64 id = p.noncommitted[k];
                  if (id == cb && cb != 1ull) c_committed = false;
                  if (id == db && db != 1ull) d_committed = false;
                }
                if (cb == p.txn_id && p.txn_id != 1ull) vis = db != p.txn_id;
                else if (!c_committed) vis = false;
                else if (cb > p.snapshot_id) vis = false;
                else if (db == ~0ull) vis = true;
                else if (db == p.txn_id && p.txn_id != 1ull) vis = false;
                else if (!d_committed) vis = true;
                else vis = db > p.snapshot_id;
              }
              m |= (unsigned)vis << r;
            }
            actm &= m;
            break;
          }
          case FO_SELECT_DONE:
            if (!__any_sync(FULL, actm != 0)) done = true;
            break;
          case FO_GROUP: {
            unsigned mine = 0;
            u64 keys[R];
#pragma unroll
            for (int r = 0; r < R; ++r) keys[r] = 0;
            {
              int shift = 0;
              for (uint32_t k = 0; k < p.n_keys; ++k) {  // key-major: each key column's constants are read once
                i64 v[R];
                load_col(p.key_col[k], p.key_load[k], v);
                const int bits = p.key_bits[k];
                const u64 kmin = p.key_min[k];
                const bool is_str = p.key_kind[k] == KK_STR;
                const int L = p.key_strlen[k];
                const u64 mask = bits == 64 ? ~0ull : ((1ull << bits) - 1);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  const u64 f = is_str ? ((L ? (((u64)v[r] >> (64 - 8 * L)) << 3) : 0ull) | ((u64)v[r] & 7ull)) : (u64)v[r] - kmin;
                  if (p.single_wide_key) keys[r] = (u64)v[r];
                  else keys[r] |= (f & mask) << shift;
                }
                shift += bits;
              }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const bool a = (actm >> r) & 1u;
              const u64 K = a ? keys[r] : kEmptyKey;
              const bool usable = a && K != kEmptyKey;
              const unsigned peers = __match_any_sync(FULL, usable ? K : kEmptyKey);
              int sl = -1;
              if (usable && lane == __ffs(peers) - 1 && FG > 0) {  // one lane per distinct key probes the CTA's table
                uint32_t h = (uint32_t)mix64(K) & (FG - 1);
                for (uint32_t i = 0; i < FG; ++i) {
                  const u64 cur = tbl[h];
                  if (cur == K) { sl = (int)h; break; }
                  if (cur == kEmptyKey) {
                    const u64 old = atomicCAS(&tbl[h], kEmptyKey, K);
                    if (old == kEmptyKey || old == K) { sl = (int)h; break; }
                  }
                  h = (h + 1) & (FG - 1);
                }
              }
              sl = __shfl_sync(FULL, sl, __ffs(peers) - 1);
              slot[r] = usable ? sl : -1;
              if (usable && sl >= 0) mine |= 1u << sl;
              if (a && !(usable && sl >= 0)) negm |= 1u << r;
            }
            present = __reduce_or_sync(FULL, mine);
            has_slow = __any_sync(FULL, negm != 0);
            break;
          }

          case FO_LD_COL: case FO_LD_LIT: case FO_LD_TMP: break;  // the pre-load above is the whole instruction
          case FO_ST_TMP: {
            i64* t = tmp_base + (size_t)in.a * T;
#pragma unroll
            for (int r = 0; r < R; ++r) t[r * NC + tid] = acc[r];
            break;
          }
          case FO_OP_COL: {
            i64 v[R];
            load_col(in.c, in.b, v);
            binop(in.a, v);
            break;
          }
          case FO_OP_LIT: {
            i64 v[R];
            const i64 l = (i64)p.lits[in.c].lo;
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = l;
            binop(in.a, v);
            break;
          }
          case FO_OP_TMP: {
            i64 v[R];
            const i64* t = tmp_base + (size_t)in.b * T;
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = t[r * NC + tid];
            binop(in.a, v);
            break;
          }
          case FO_DIVR: {
            if (in.b == 2) {  // 0 <= x < 2^32
              const unsigned d = (unsigned)kPow10U64[in.a], half = d / 2;
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const unsigned x = (unsigned)acc[r];
                const unsigned q = x / d;
                acc[r] = (i64)(q + ((x - q * d) >= half ? 1u : 0u));
              }
            } else if (in.b == 1) {  // x >= 0
              const u64 d = kPow10U64[in.a], half = d / 2;
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const u64 x = (u64)acc[r];
                const u64 q = x / d;
                acc[r] = (i64)(q + ((x - q * d) >= half ? 1ull : 0ull));
              }
            } else {
#pragma unroll
              for (int r = 0; r < R; ++r) acc[r] = div_pow10_round<i64>(acc[r], in.a);
            }
            break;
          }
          case FO_MULP: {
            const i64 m = pow10_i64(in.a);
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] *= m;
            break;
          }
          case FO_I2F:
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = f_bits(__ll2double_rn(acc[r]));
            break;
          case FO_D2F: {
            const double den = __longlong_as_double((i64)p.lits[in.c].lo);
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = f_bits(__ll2double_rn(acc[r]) / den);
            break;
          }

          // ------------------------------------------------------------ aggregates: group-major over the slots present
          // in this warp; a lane first folds its own R rows, then the warp reduces once per group
          case FO_COUNT_STAR: case FO_COUNT: {
            if (has_slow) slow_rows(in.op, in.a, negm, in.c);
            const unsigned okm = actm & ~negm;
            for (unsigned gm = present; gm; gm &= gm - 1) {
              const int g = __ffs(gm) - 1;
              unsigned c = 0;
#pragma unroll
              for (int r = 0; r < R; ++r) c += ((okm >> r) & 1u) && slot[r] == g;
              c = __reduce_add_sync(FULL, c);
              if (lane == 0) my_acc[(uint32_t)g * NFW + in.b] += c;
            }
            break;
          }
          case FO_FIRSTROW: case FO_FIRSTVALID: case FO_FIRSTNAN: {
            unsigned setm = actm;
            if (in.op == FO_FIRSTNAN) {
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const double d = f_as_f64(acc[r]);
                if (d == d) setm &= ~(1u << r);
              }
            }
            if (has_slow) slow_rows(in.op, in.a, setm & negm, in.c);
            setm &= ~negm;
            for (unsigned gm = present; gm; gm &= gm - 1) {
              const int g = __ffs(gm) - 1;
              unsigned off = 0xffffffffu;
#pragma unroll
              for (int r = R - 1; r >= 0; --r)
                if (((setm >> r) & 1u) && slot[r] == g) off = (unsigned)(r * NC + tid);
              off = __reduce_min_sync(FULL, off);
              if (lane == 0 && off != 0xffffffffu) {
                u64* a = &my_acc[(uint32_t)g * NFW + in.b];
                const u64 cand = row0 + off;
                if (cand < *a) *a = cand;
              }
            }
            break;
          }
          case FO_SUM: {
            const int limbs = in.a & 3;  // 0: check each value
            unsigned fastm = actm & ~negm;
            if (limbs == 0) {
              unsigned bigm = 0;
#pragma unroll
              for (int r = 0; r < R; ++r)
                if (!(acc[r] < ((i64)1 << 40) && acc[r] > -((i64)1 << 40))) bigm |= 1u << r;
              bigm &= actm;
              if (__any_sync(FULL, (bigm | negm) != 0)) slow_rows(FO_SUM, in.a, bigm | negm, in.c);
              fastm &= ~bigm;
            } else if (has_slow) {
              slow_rows(FO_SUM, in.a, negm, in.c);
            }
            for (unsigned gm = present; gm; gm &= gm - 1) {
              const int g = __ffs(gm) - 1;
              i64 x = 0;
#pragma unroll
              for (int r = 0; r < R; ++r)
                if (((fastm >> r) & 1u) && slot[r] == g) x += acc[r];
              i64 tot;
              if (limbs == 1) {
                tot = (i64)__reduce_add_sync(FULL, (int)x);  // |x| < 2^24 per lane
              } else if (limbs == 2) {
                const unsigned lo24 = __reduce_add_sync(FULL, (unsigned)((u64)x & 0xffffffull));
                const int hi = __reduce_add_sync(FULL, (int)(x >> 24));
                tot = ((i64)hi << 24) + (i64)lo24;
              } else {
                const unsigned lo24 = __reduce_add_sync(FULL, (unsigned)((u64)x & 0xffffffull));
                const unsigned mid = __reduce_add_sync(FULL, (unsigned)((u64)(x >> 24) & 0xffffffull));
                const int top = __reduce_add_sync(FULL, (int)(x >> 48));
                tot = ((i64)top << 48) + ((i64)mid << 24) + (i64)lo24;
              }
              if (lane == 0) {
                u64* a = &my_acc[(uint32_t)g * NFW + in.b];
                *a = (u64)((i64)*a + tot);
              }
            }
            break;
          }
          case FO_FSUM: {
            if (has_slow) slow_rows(FO_FSUM, in.a, negm, in.c);
            const unsigned okm = actm & ~negm;
            for (unsigned gm = present; gm; gm &= gm - 1) {
              const int g = __ffs(gm) - 1;
              double x = 0.0;
#pragma unroll
              for (int r = 0; r < R; ++r)
                if (((okm >> r) & 1u) && slot[r] == g) x += f_as_f64(acc[r]);
              for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
              if (lane == 0) {
                u64* a = &my_acc[(uint32_t)g * NFW + in.b];
                *a = (u64)f_bits(f_as_f64((i64)*a) + x);
              }
            }
            break;
          }
          case FO_MIN_I: case FO_MAX_I: case FO_MIN_F: case FO_MAX_F: {
            const bool is_min = in.op == FO_MIN_I || in.op == FO_MIN_F;
            const bool is_f = in.op == FO_MIN_F || in.op == FO_MAX_F;
            unsigned setm = actm;
            u64 e[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              if (is_f) {
                const double d = f_as_f64(acc[r]);
                if (d != d) setm &= ~(1u << r);  // NaN never replaces a number (llkv-aggregate/src/lib.rs:1309-1331)
                e[r] = enc_f64(d);
              } else e[r] = enc_i64(acc[r]);
            }
            if (has_slow) slow_rows(in.op, in.a, setm & negm, in.c);
            setm &= ~negm;
            for (unsigned gm = present; gm; gm &= gm - 1) {
              const int g = __ffs(gm) - 1;
              u64 x = is_min ? ~0ull : 0ull;
#pragma unroll
              for (int r = 0; r < R; ++r)
                if (((setm >> r) & 1u) && slot[r] == g) x = is_min ? (e[r] < x ? e[r] : x) : (e[r] > x ? e[r] : x);
              for (int o = 16; o; o >>= 1) {
                const u64 y = __shfl_xor_sync(FULL, x, o);
                x = is_min ? (y < x ? y : x) : (y > x ? y : x);
              }
              if (lane == 0) {
                u64* a = &my_acc[(uint32_t)g * NFW + in.b];
                if (is_min ? x < *a : x > *a) *a = x;
              }
            }
            break;
          }
          case FO_END: done = true; break;
          default: errbits |= FLAG_BAD_PLAN; done = true; break;
        }
        if (done) break;
        in = p.fcode[++pc];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);  // this warp is done with the stage
    }
  }

  // ---------------------------------------------------------------- fold the warps' rows into the global group table
  __syncthreads();
  for (uint32_t i = tid; i < FG * NFW; i += blockDim.x) {
    const uint32_t g = i / NFW, w = i % NFW;
    const u64 K = tbl[g];
    if (p.n_keys != 0 && K == kEmptyKey) continue;
    const FastWord fw = p.fast[w];
    if (fw.kind == FK_SKIP) continue;
    const u64 gslot = p.n_keys != 0 ? global_slot(p, K, false, errbits) : 0;
    u64* grow = &p.gwords[gslot * p.n_gwords];
    switch (fw.kind) {
      case FK_COUNT: {
        u64 s = 0;
        for (int cw = 0; cw < n_cwarps; ++cw) s += wacc[((size_t)cw * FG + g) * NFW + w];
        if (s) atomicAdd(&grow[fw.gword], s);
        break;
      }
      case FK_SUM_I64: case FK_SUM_I128: {
        i128 t = 0;
        for (int cw = 0; cw < n_cwarps; ++cw) t += (i128)(i64)wacc[((size_t)cw * FG + g) * NFW + w];
        if (t != 0) {
          if (fw.kind == FK_SUM_I64) gadd_sum_i64(&grow[fw.gword], t);
          else gadd_sum_i128(&grow[fw.gword], t);
        }
        break;
      }
      case FK_FSUM: {
        double s = 0.0;
        for (int cw = 0; cw < n_cwarps; ++cw) s += __longlong_as_double((i64)wacc[((size_t)cw * FG + g) * NFW + w]);
        atomicAdd(reinterpret_cast<double*>(&grow[fw.gword]), s);
        break;
      }
      case FK_MIN: {
        u64 s = ~0ull;
        for (int cw = 0; cw < n_cwarps; ++cw) s = min(s, wacc[((size_t)cw * FG + g) * NFW + w]);
        atomicMin(&grow[fw.gword], s);
        break;
      }
      case FK_MAX: {
        u64 s = 0ull;
        for (int cw = 0; cw < n_cwarps; ++cw) s = max(s, wacc[((size_t)cw * FG + g) * NFW + w]);
        atomicMax(&grow[fw.gword], s);
        break;
      }
      default: errbits |= FLAG_BAD_PLAN; break;
    }
  }
  if (errbits) atomicOr(p.flags, errbits);
}

cudaError_t launch_fast(const Plan* dplan, int rows_per_thread, uint32_t grid, uint32_t consumer_threads, uint32_t smem, cudaStream_t stream) {
  const uint32_t block = consumer_threads + 32;
#define LLKV_LAUNCH_FAST(RR)                                                                                                  \
  do {                                                                                                                        \
    cudaError_t e = cudaFuncSetAttribute(fast_scan_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    if (e != cudaSuccess) return e;                                                                                           \
    fast_scan_kernel<RR><<<grid, block, smem, stream>>>(dplan);                                                               \
    return cudaGetLastError();                                                                                                \
  } while (0)
  if (rows_per_thread == 8) LLKV_LAUNCH_FAST(8);
  if (rows_per_thread == 4) LLKV_LAUNCH_FAST(4);
  if (rows_per_thread == 2) LLKV_LAUNCH_FAST(2);
  LLKV_LAUNCH_FAST(1);
#undef LLKV_LAUNCH_FAST
}

}  // namespace llkv
