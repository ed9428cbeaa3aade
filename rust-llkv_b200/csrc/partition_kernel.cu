// partition_kernel.cu — pass 2 of the partitioned high-cardinality GROUP BY (pass 1 is LeanTile::scatter() in
// lean_kernel.cuh).  The scan left (key, row id, operands) tuples in 2^bits hash partitions, each covering a contiguous
// slice of the global open-addressing table that fits in L2.  This kernel walks the partitions in order — every CTA is on
// the same one or two partitions at any time — so the probes and the atomics of a slice hit L2 instead of random DRAM
// sectors, while the tuples stream in once, coalesced, with evict-first loads.
//
// A thread works on kPartUnroll tuples at a time: all their fields are requested first (DRAM latency, once), then all
// first probes (L2 latency, once), then the updates (fire-and-forget REDs).  Loading a field where it is used instead
// serialises one DRAM round trip per field and tuple and was 2.5x slower.
#include "lean_kernel.cuh"

namespace llkv {

constexpr int kPartThreads = 256;

template <int NV, int kPartUnroll, int kMinBlocks>  // NV: operand fields held in registers (>= pp.n_vops)
__global__ void __launch_bounds__(kPartThreads, kMinBlocks) partition_apply_kernel(const __grid_constant__ PartPlan pp) {
  constexpr int NVR = NV ? NV : 1;
  const u64 mask = pp.gcap - 1;
  const u64 cap = pp.part_cap;
  uint32_t errbits = 0;
  // Work items = the filled chunks of all partitions, in partition order; CTA b takes items b, b + grid, ...  Every CTA
  // gets the same number of chunks (+-1) and all of them sit in the same one or two partitions at any time, which is
  // what keeps the table slice in L2.
  __shared__ uint32_t s_first[kMaxPartitions + 1];  // first item of each partition
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (uint32_t q = 0; q < pp.n_parts; ++q) {
      s_first[q] = run;
      const u64 filled = pp.cursor[q];
      const u64 n = filled < cap ? filled : cap;  // tuples past the capacity were applied by the scan itself
      run += (uint32_t)((n + pp.chunk - 1) / pp.chunk);
    }
    s_first[pp.n_parts] = run;
  }
  __syncthreads();
  const uint32_t items = s_first[pp.n_parts];
  uint32_t q = 0;
  for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
    while (item >= s_first[q + 1]) ++q;
    const uint32_t c = item - s_first[q];
    const u64 filled = pp.cursor[q];
    const u64 n = filled < cap ? filled : cap;
    const u64 start = (u64)c * pp.chunk;
    const u64 end = n < start + pp.chunk ? n : start + pp.chunk;
    const u64* base = pp.tuples + (u64)q * pp.n_fields * cap;
    // The loop is warp-uniform and every phase reconverges (__syncwarp / __any_sync): a warp that stays split after the
    // probe loop issues each coalesced tuple load once per fragment (measured: 3 fragments, 6x the tuple bytes from L2).
    for (u64 b0 = start; b0 < end; b0 += (u64)kPartThreads * kPartUnroll) {
      __syncwarp();
      const u64 i0 = b0 + threadIdx.x;
      u64 K[kPartUnroll], row[kPartUnroll], fv[kPartUnroll][NVR], h[kPartUnroll], cur[kPartUnroll];
      bool on[kPartUnroll];
#pragma unroll
      for (int u = 0; u < kPartUnroll; ++u) {
        const u64 i = i0 + (u64)u * kPartThreads;
        on[u] = i < end;
        K[u] = on[u] ? __ldcs(base + i) : kEmptyKey;
        row[u] = on[u] ? __ldcs(base + cap + i) : 0ull;
#pragma unroll
        for (int j = 0; j < NV; ++j) fv[u][j] = (on[u] && j < (int)pp.n_vops) ? __ldcs(base + (u64)(2 + j) * cap + i) : 0ull;
      }
      unsigned need = 0;  // bit u: tuple u still looks for its slot
#pragma unroll
      for (int u = 0; u < kPartUnroll; ++u) {
        const bool probe = on[u] && K[u] != kEmptyKey && pp.n_keys != 0;
        h[u] = probe ? (mix64(K[u]) & mask) : (pp.n_keys ? pp.gcap : 0ull);  // the reserved key value has its own row
        cur[u] = probe ? __ldcg(&pp.gkeys[h[u]]) : 0ull;
        need |= (unsigned)probe << u;
      }
      // one probe step of every unresolved tuple per round: the next slots of a thread's tuples are requested together
      u64 rounds = 0;
      while (__any_sync(LLKV_FULL, need != 0)) {
#pragma unroll
        for (int u = 0; u < kPartUnroll; ++u) {
          if (!((need >> u) & 1u)) continue;
          const u64 cc = cur[u];
          bool done = cc == K[u];
          if (!done && cc == kEmptyKey) {
            const u64 old = atomicCAS(&pp.gkeys[h[u]], kEmptyKey, K[u]);
            done = old == kEmptyKey || old == K[u];
          }
          if (done) need &= ~(1u << u);
          else {
            h[u] = (h[u] + 1) & mask;
            cur[u] = __ldcg(&pp.gkeys[h[u]]);
          }
        }
        if (++rounds > mask && need) {  // every slot seen: the table is full
          errbits |= FLAG_TABLE_FULL;
#pragma unroll
          for (int u = 0; u < kPartUnroll; ++u)
            if ((need >> u) & 1u) h[u] = pp.gcap;
          need = 0;
        }
      }
      __syncwarp();
      // aggregate-major: the words of one group row share a sector, and back-to-back atomics on one sector queue up in
      // the L2 atomic unit; other tuples' updates go in between
      for (uint32_t k = 0; k < pp.n_nops; ++k) {
        const PartOp op = pp.nops[k];
        const bool is_count = op.op == FO_COUNT_STAR || op.op == FO_COUNT;
#pragma unroll
        for (int u = 0; u < kPartUnroll; ++u) {
          if (!on[u]) continue;
          u64* w = pp.gwords + h[u] * pp.n_gwords + op.gword;
          if (is_count) atomicAdd(w, 1ull);
          else atomicMin(w, row[u]);  // FO_FIRSTROW / FO_FIRSTVALID
        }
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (j >= (int)pp.n_vops) break;
        const PartOp op = pp.vops[j];
        const bool plain_sum = op.op == FO_SUM && !(op.flags & 0x80);
#pragma unroll
        for (int u = 0; u < kPartUnroll; ++u) {
          if (!on[u]) continue;
          u64* w = pp.gwords + h[u] * pp.n_gwords + op.gword;
          const u64 f = fv[u][j];
          if (plain_sum) {  // gadd_sum_i64 without the second atomic when it adds nothing
            atomicAdd(&w[0], f & 0xffffffffull);
            const u64 hi = (u64)((i64)f >> 32);
            if (hi) atomicAdd(&w[1], hi);
          } else lean_apply_field(w, op.op, op.flags, f, row[u]);
        }
      }
    }
  }
  if (errbits) atomicOr(pp.flags, errbits);
}

cudaError_t launch_partition_apply(const PartPlan& plan, uint32_t grid, cudaStream_t stream) {
  const uint32_t nv = plan.n_vops;
  // grid = `sms` x the kernel's resident CTAs per SM (the caller passes grid = SM count): all CTAs are resident at once.
  // Four tuples per thread and 1024 threads per SM measured best (tools/exp_part.py history in profiles/r01_summary.md:
  // two tuples x 1536 threads 7.4 ms, eight x 512 6.8 ms, four x 1024 6.5 ms per 200 M rows).
#define LLKV_PART_LAUNCH(NV, U, B)                                                \
  do {                                                                            \
    partition_apply_kernel<NV, U, B><<<grid * B, kPartThreads, 0, stream>>>(plan); \
    return cudaGetLastError();                                                    \
  } while (0)
  if (nv == 0) LLKV_PART_LAUNCH(0, 4, 4);
  if (nv == 1) LLKV_PART_LAUNCH(1, 4, 4);
  if (nv == 2) LLKV_PART_LAUNCH(2, 4, 4);
  if (nv <= 4) LLKV_PART_LAUNCH(4, 4, 2);
  if (nv <= kMaxPartOperands) LLKV_PART_LAUNCH(kMaxPartOperands, 4, 2);
#undef LLKV_PART_LAUNCH
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------------------------------
// Pass 2 of the packed form (LeanTile::scatter_packed): every partition holds the tuples of at most `slots` groups or so,
// and one CTA aggregates a whole partition in shared memory — 32-bit shared-memory atomics only (64-bit shared atomics are
// CAS loops on this architecture): a row count, the first row (MIN) and, per SUM operand, a (low, carry) pair of u32 that
// is exact for operands below 2^32.  Dense integer keys index the slots directly (a partition is a key range); other keys
// go through an open-addressing table of the partition's keys in shared memory.  When the partition is done each group
// is written to the global table once: the per-row global atomics of the first form (1.2 probes + 3 REDs per row, bound by
// L2 request rate) become one insert per group and launch.
constexpr int kFoldThreads = 1024;

template <int NOPS, bool DENSE>
__global__ void __launch_bounds__(kFoldThreads, 1) partition_fold_kernel(const __grid_constant__ FoldPlan fp) {
  extern __shared__ __align__(16) unsigned char fold_smem[];
  const uint32_t S = fp.slots, smask = S - 1;
  uint32_t* const s_rows = reinterpret_cast<uint32_t*>(fold_smem);   // [S]
  uint32_t* const s_first = s_rows + S;                               // [S]
  uint32_t* const s_lo = s_first + S;                                 // [NOPS][S]
  uint32_t* const s_hi = s_lo + (size_t)(NOPS ? NOPS : 1) * S;        // [NOPS][S]
  u64* const s_keys = reinterpret_cast<u64*>(s_hi + (size_t)(NOPS ? NOPS : 1) * S);  // [S] (hashed form only)
  const uint32_t tid = threadIdx.x;
  const u64 kmask = (1ull << fp.key_bits) - 1, rmask = (1ull << fp.row_bits) - 1;
  const u64 cap = fp.part_cap;
  uint32_t errbits = 0;
  uint32_t opshift[NOPS ? NOPS : 1];
  u64 opmask[NOPS ? NOPS : 1];
  {
    uint32_t sh = fp.key_bits + fp.row_bits;
#pragma unroll
    for (int j = 0; j < NOPS; ++j) {
      opshift[j] = sh;
      opmask[j] = (1ull << fp.op_bits[j]) - 1;
      sh += fp.op_bits[j];
    }
  }
  for (uint32_t q = blockIdx.x; q < fp.n_parts; q += gridDim.x) {
    const u64 filled = fp.cursor[q];
    const u64 n = filled < cap ? filled : cap;  // tuples past the capacity were applied by the scan itself
    if (n == 0) continue;
    for (uint32_t s = tid; s < S; s += kFoldThreads) {
      s_rows[s] = 0;
      s_first[s] = 0xffffffffu;
#pragma unroll
      for (int j = 0; j < NOPS; ++j) {
        s_lo[j * S + s] = 0;
        s_hi[j * S + s] = 0;
      }
      if (!DENSE) s_keys[s] = kEmptyKey;
    }
    __syncthreads();
    const u64* base = fp.tuples + (u64)q * cap;
    constexpr int U = 4;  // tuples per thread and round; the next round's loads are in flight while this one is folded
    u64 nxt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const u64 i = (u64)u * kFoldThreads + tid;
      nxt[u] = i < n ? __ldcs(base + i) : ~0ull;
    }
    for (u64 b0 = 0; b0 < n; b0 += (u64)kFoldThreads * U) {
      u64 t[U];
#pragma unroll
      for (int u = 0; u < U; ++u) t[u] = nxt[u];
      const u64 b1 = b0 + (u64)kFoldThreads * U;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const u64 i = b1 + (u64)u * kFoldThreads + tid;
        nxt[u] = i < n ? __ldcs(base + i) : ~0ull;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (b0 + (u64)u * kFoldThreads + tid >= n) continue;
        const u64 K = t[u] & kmask;
        uint32_t slot;
        if (DENSE) {
          slot = (uint32_t)K & smask;
        } else {
          slot = (uint32_t)(mix64(K) >> 40) & smask;  // (high bits: the low bits chose the partition)
          uint32_t tries = 0;
          for (;;) {
            const u64 cur = s_keys[slot];
            if (cur == K) break;
            if (cur == kEmptyKey) {
              const u64 old = atomicCAS(&s_keys[slot], kEmptyKey, K);
              if (old == kEmptyKey || old == K) break;
            }
            slot = (slot + 1) & smask;
            if (++tries > smask) { slot = ~0u; break; }
          }
        }
        const uint32_t rel = (uint32_t)((t[u] >> fp.key_bits) & rmask);
        if (slot == ~0u) {  // more distinct keys than slots in this partition (skew): this tuple goes to the table directly
          const u64 gs = lean_global_slot(fp.gkeys, fp.gcap, fp.n_keys, K, errbits);
          u64* w = fp.gwords + gs * fp.n_gwords;
          for (uint32_t c = 0; c < fp.n_counts; ++c) atomicAdd(&w[fp.count_gword[c]], 1ull);
          if (fp.first_gword != ~0u) atomicMin(&w[fp.first_gword], fp.row_base + rel);
#pragma unroll
          for (int j = 0; j < NOPS; ++j) {
            const u64 v = (t[u] >> opshift[j]) & opmask[j];
            if (fp.op_wide[j]) gadd_sum_i128(&w[fp.op_gword[j]], (i128)v);
            else gadd_sum_i64(&w[fp.op_gword[j]], (i128)v);
          }
          continue;
        }
        atomicAdd(&s_rows[slot], 1u);
        atomicMin(&s_first[slot], rel);
#pragma unroll
        for (int j = 0; j < NOPS; ++j) {
          const uint32_t v = (uint32_t)((t[u] >> opshift[j]) & opmask[j]);
          const uint32_t old = atomicAdd(&s_lo[j * S + slot], v);
          if (old + v < old) atomicAdd(&s_hi[j * S + slot], 1u);
        }
      }
    }
    __syncthreads();
    // every group of the partition, once
    for (uint32_t s = tid; s < S; s += kFoldThreads) {
      const uint32_t rows = s_rows[s];
      if (!rows) continue;
      const u64 K = DENSE ? (((u64)q << fp.part_shift) | s) : s_keys[s];
      const u64 gs = lean_global_slot(fp.gkeys, fp.gcap, fp.n_keys, K, errbits);
      u64* w = fp.gwords + gs * fp.n_gwords;
      for (uint32_t c = 0; c < fp.n_counts; ++c) atomicAdd(&w[fp.count_gword[c]], (u64)rows);
      if (fp.first_gword != ~0u) atomicMin(&w[fp.first_gword], fp.row_base + s_first[s]);
#pragma unroll
      for (int j = 0; j < NOPS; ++j) {
        const u64 total = ((u64)s_hi[j * S + s] << 32) | s_lo[j * S + s];
        if (!total) continue;
        if (fp.op_wide[j]) gadd_sum_i128(&w[fp.op_gword[j]], (i128)total);
        else gadd_sum_i64(&w[fp.op_gword[j]], (i128)total);
      }
    }
    __syncthreads();
  }
  if (errbits) atomicOr(fp.flags, errbits);
}

// (the kernel is instantiated for 0, 1, 2 and 4 operands: three use the layout of four)
uint32_t fold_smem_bytes(uint32_t slots, uint32_t n_ops, bool dense) {
  const uint32_t ops = n_ops <= 1 ? 1 : (n_ops == 2 ? 2 : 4);
  return slots * (8u + 8u * ops + (dense ? 0u : 8u));
}

cudaError_t launch_partition_fold(const FoldPlan& plan, uint32_t grid, cudaStream_t stream) {
  const uint32_t smem = fold_smem_bytes(plan.slots, plan.n_ops, plan.dense != 0);
#define LLKV_FOLD_LAUNCH(NOPS, D)                                                                                                   \
  do {                                                                                                                              \
    cudaError_t e = cudaFuncSetAttribute(partition_fold_kernel<NOPS, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    if (e != cudaSuccess) return e;                                                                                                 \
    partition_fold_kernel<NOPS, D><<<grid, kFoldThreads, smem, stream>>>(plan);                                                    \
    return cudaGetLastError();                                                                                                      \
  } while (0)
  if (plan.dense) {
    if (plan.n_ops == 0) LLKV_FOLD_LAUNCH(0, true);
    if (plan.n_ops == 1) LLKV_FOLD_LAUNCH(1, true);
    if (plan.n_ops == 2) LLKV_FOLD_LAUNCH(2, true);
    if (plan.n_ops <= 4) LLKV_FOLD_LAUNCH(4, true);
  } else {
    if (plan.n_ops == 0) LLKV_FOLD_LAUNCH(0, false);
    if (plan.n_ops == 1) LLKV_FOLD_LAUNCH(1, false);
    if (plan.n_ops == 2) LLKV_FOLD_LAUNCH(2, false);
    if (plan.n_ops <= 4) LLKV_FOLD_LAUNCH(4, false);
  }
#undef LLKV_FOLD_LAUNCH
  return cudaErrorInvalidValue;
}

}  // namespace llkv
