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
constexpr int kPartUnroll = 4;

template <int NV>  // operand fields held in registers (>= pp.n_vops)
__global__ void __launch_bounds__(kPartThreads, NV <= 2 ? 4 : 2) partition_apply_kernel(const __grid_constant__ PartPlan pp) {
  constexpr int NVR = NV ? NV : 1;
  const u64 mask = pp.gcap - 1;
  const u64 cap = pp.part_cap;
  uint32_t errbits = 0;
  const uint32_t items = pp.n_parts * pp.chunks_per_part;
  for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
    const uint32_t q = item / pp.chunks_per_part, c = item % pp.chunks_per_part;
    const u64 filled = pp.cursor[q];
    const u64 n = filled < cap ? filled : cap;  // tuples past the capacity were applied by the scan itself
    const u64 start = (u64)c * pp.chunk;
    if (start >= n) continue;
    const u64 end = n < start + pp.chunk ? n : start + pp.chunk;
    const u64* base = pp.tuples + (u64)q * pp.n_fields * cap;
    for (u64 i0 = start + threadIdx.x; i0 < end; i0 += (u64)kPartThreads * kPartUnroll) {
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
#pragma unroll
      for (int u = 0; u < kPartUnroll; ++u) {
        h[u] = mix64(K[u]) & mask;
        cur[u] = (K[u] != kEmptyKey && pp.n_keys) ? __ldcg(&pp.gkeys[h[u]]) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < kPartUnroll; ++u) {
        if (!on[u]) continue;
        u64 gs = pp.gcap;  // the reserved key value has its own row
        if (pp.n_keys == 0) gs = 0;
        else if (K[u] != kEmptyKey) {
          u64 hh = h[u], cc = cur[u], probes = 0;
          while (true) {
            if (cc == K[u]) break;
            if (cc == kEmptyKey) {
              const u64 old = atomicCAS(&pp.gkeys[hh], kEmptyKey, K[u]);
              if (old == kEmptyKey || old == K[u]) break;
            }
            if (++probes > mask) {
              errbits |= FLAG_TABLE_FULL;
              hh = pp.gcap;
              break;
            }
            hh = (hh + 1) & mask;
            cc = __ldcg(&pp.gkeys[hh]);
          }
          gs = hh;
        }
        h[u] = gs;
      }
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
  if (nv == 0) partition_apply_kernel<0><<<grid, kPartThreads, 0, stream>>>(plan);
  else if (nv == 1) partition_apply_kernel<1><<<grid, kPartThreads, 0, stream>>>(plan);
  else if (nv == 2) partition_apply_kernel<2><<<grid, kPartThreads, 0, stream>>>(plan);
  else if (nv <= 4) partition_apply_kernel<4><<<grid, kPartThreads, 0, stream>>>(plan);
  else if (nv <= kMaxPartOperands) partition_apply_kernel<kMaxPartOperands><<<grid, kPartThreads, 0, stream>>>(plan);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace llkv
