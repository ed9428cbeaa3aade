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

}  // namespace llkv
