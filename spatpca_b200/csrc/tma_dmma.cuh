// Building blocks of the TMA-fed FP64 DMMA matrix stream (sm_100a):  U (p x NC) = Omega_u (p x p) * V (p x NC),
// NC <= 32 vectors -- the p x p . p x K contraction of the north-star, batched over the folds of a stage
// (NC = folds x K).  With 25-32 vectors the product is no longer HBM-bound (2 p^2 NC flop over 8 p^2 bytes =
// 6-8 flop/B, at the FP64 ridge of B200), so it runs on the FP64 tensor path: mma.sync m8n8k4 (DMMA), operands
// staged in shared memory by TMA (cp.async.bulk.tensor + mbarrier, SWIZZLE_128B), accumulators in registers.
// There is no FP64 tcgen05 kind; DMMA is the FP64 tensor instruction of this part.
//
// Decomposition (no split-K, no atomics, bitwise reproducible): CTA b owns the J = 8 NT output rows
// [b J, (b + 1) J) for the whole kernel.  By symmetry of Omega_u it streams the COLUMNS J of the matrix (contiguous
// 128-byte row segments per column) over all p rows in chunks of CR rows:
//     U'[:, J] (NC x J)  +=  V'[:, chunk] (NC x CR)  *  Omega_u[chunk, J] (CR x J)
// i.e. V' is the MMA's A operand (M = vectors), Omega_u the B operand (N = matrix columns = output rows).
// Consumer warp h owns the h-th group of 8 rows of every 64-row chunk, for all vector tiles and all NT column tiles
// (2 MT NT accumulator registers per lane); the 8 partial sums of a tile are added in warp order through shared
// memory at the end of a pass.
//
// Shared-memory layout and bank conflicts.  A TMA box is 16 rows x C columns; with SWIZZLE_128B column c is one
// 128-byte line whose 16-byte chunk x sits at chunk position x ^ (c % 8).  A lane (g = lane / 4, t = lane % 4)
// of an m8n8k4 needs B[k = t][n = g]: it loads the row PAIR (2t, 2t + 1) of its column with one 16-byte load
// and uses the two halves for two MMAs -- the contraction index may be enumerated in any order as long as both
// operands agree, so MMA step s of a row group takes rows {s, 2 + s, 4 + s, 6 + s}.  The 8 lanes of a
// quarter-warp then read chunks t (4 values) of two columns; mapping MMA column slot g to tile column
// perm(g) = 4 (g % 2) + g / 2 puts the two columns 4 apart, the swizzle flips chunk bit 2 between them, and the
// 8 loads hit 8 distinct chunk positions: conflict-free.  Outputs are un-permuted when they are stored.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace spb {
namespace tmx {

constexpr int kConsumerWarps = 8;

// MT = vector tiles (of 8): 1, 2 or 4.  Every consumer warp takes ONE group of 8 rows of each 64-row chunk and ALL
// MT x NT tile pairs of it, so a matrix fragment is loaded once per CTA (not once per vector tile) and a warp has
// MT * NT independent accumulators in flight.  With MT = 4 that is up to 80 FP64 accumulators = 160 registers per
// lane: those instances run 8 warps (255 registers each) and lane 0 of warp 0 doubles as the TMA producer; the
// MT = 1 / 2 instances are HBM-bound and keep a dedicated ninth producer warp, which never stalls behind MMA work
// (the register file is allocated in units of four warps, so nine warps cap a thread at 168 registers).
template <int MT>
struct Shape {
  static constexpr int CR = 64;                          // rows per chunk
  static constexpr int NBX = CR / 16;                    // TMA boxes (16 rows) per chunk and operand
  static constexpr int NCP = 8 * MT;                     // vectors, padded
  static constexpr bool kDedicatedProducer = (MT < 4);
  // MT = 4 (FP64-tensor-bound): 12 consumer warps = 4 row-group pairs x 3 column-tile sets, 3 warps per scheduler
  // (ncu on the 8-warp version: DMMA pipe 62 % busy, 38 % of the stall samples `wait`, i.e. too few warps to cover
  // the fixed issue latency of back-to-back DMMAs); lane 0 of warp 0 doubles as the TMA producer.
  static constexpr int kConsumers = (MT == 4) ? 12 : kConsumerWarps;
  static constexpr int kThreads = 32 * (kConsumers + (kDedicatedProducer ? 1 : 0));
  static constexpr int kMaxNT = (MT == 4) ? 15 : 16;     // column tiles (of 8) per CTA
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "SPB_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra SPB_MBAR_DONE;\n"
      "bra SPB_MBAR_WAIT;\n"
      "SPB_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// 2-D tile load: box of the tensor map at (c0 = row, c1 = column) -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// D (8 x 8) += A (8 x 4, row) * B (4 x 8, col), FP64 tensor core
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ int perm8(int g) { return ((g & 1) << 2) | (g >> 1); }


struct StreamGeom {
  int p;            // matrix order
  int NT;           // column tiles per CTA
  int J;            // 8 * NT: output rows (= matrix columns) owned by a CTA
  int nchunks;      // ceil(p / CR)
  int NS;           // pipeline stages
  int D;            // chunks the integrated producer runs ahead (MT = 4 instances)
  uint32_t stageA;  // bytes of one matrix stage: CR * J * 8
  uint32_t stageV;  // bytes of one vector stage: CR * NCP * 8
};

// One pass  us[v][j] = sum_k V[k, v] * Omega_u[k, j0 + j]  for this CTA's J columns (all threads of the CTA call it).
// `it` is the CTA's running chunk counter (mbarrier phases continue across passes); us: [NCP][J] doubles, which MAY
// alias the stage ring (it is written after the last chunk has been consumed; the pass starts with a proxy fence so
// that the generic-proxy accesses of the previous pass are ordered before this pass's TMA writes).
// NTM: compile-time bound of the column tiles per CTA (register arrays are unrolled to it).
template <int MT, int NTM>
__device__ __noinline__ void dmma_stream_pass(const CUtensorMap* mapA, const CUtensorMap* mapV, const StreamGeom& G,
                                                 int j0, unsigned char* sA, unsigned char* sV, uint64_t* full,
                                                 uint64_t* empty, unsigned int& it, double* us) {
  using S = Shape<MT>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fence_proxy_async();
  __syncthreads();
  // chunk `c` of this pass -> its ring slot (waits until every consumer warp has released the slot's previous use)
  auto issue = [&](int c) {
    const unsigned int pit = it + (unsigned int)c;
    const int stage = pit % G.NS;
    const uint32_t phase = (pit / G.NS) & 1;
    mbar_wait(&empty[stage], phase ^ 1);
    mbar_expect_tx(&full[stage], G.stageA + G.stageV);
#pragma unroll
    for (int b = 0; b < S::NBX; ++b) {
      const int row = c * S::CR + 16 * b;
      tma_load_2d(sA + (size_t)stage * G.stageA + (size_t)b * 128 * G.J, mapA, row, j0, &full[stage]);
      tma_load_2d(sV + (size_t)stage * G.stageV + (size_t)b * 128 * S::NCP, mapV, row, 0, &full[stage]);
    }
  };
  if (S::kDedicatedProducer && warp == kConsumerWarps) {
    if (lane == 0)
      for (int c = 0; c < G.nchunks; ++c) issue(c);
  } else {
    // integrated producer: D chunks ahead (G.D <= NS - 1; with D = NS - 1 the slot refilled before chunk c was last
    // used by chunk c - 1)
    const bool producer = !S::kDedicatedProducer && threadIdx.x == 0;
    const int D = G.D;
    if (producer) {
      const int pre = G.nchunks < D ? G.nchunks : D;
      for (int c = 0; c < pre; ++c) issue(c);
    }
    const int g = lane >> 2, t = lane & 3;
    const int pg = perm8(g);
    if constexpr (MT == 4) {
      // warp = (row-group pair hq, column-tile set ns): row groups 2 hq and 2 hq + 1 of every chunk (= the two halves
      // of TMA box hq), column tiles [ns * NTM, ns * NTM + NTM) of the CTA's NT, all 4 vector tiles
      const int ns = warp % 3, hq = warp / 3;
      const int nt0 = ns * NTM;
      const uint32_t swz0 = (uint32_t)((t ^ pg) << 4), swz1 = (uint32_t)(((4 | t) ^ pg) << 4);
      const uint32_t baseA = (uint32_t)hq * 128u * (uint32_t)S::NCP + (uint32_t)pg * 128u;
      const uint32_t baseB = (uint32_t)hq * 128u * (uint32_t)G.J + (uint32_t)(nt0 * 8 + pg) * 128u;
      double acc[4][NTM][2];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int q = 0; q < NTM; ++q) { acc[mt][q][0] = 0.0; acc[mt][q][1] = 0.0; }
      unsigned int cit = it;
      for (int c = 0; c < G.nchunks; ++c, ++cit) {
        if (producer && c + D < G.nchunks) issue(c + D);
        __syncwarp();
        const int stage = cit % G.NS;
        const uint32_t phase = (cit / G.NS) & 1;
        mbar_wait(&full[stage], phase);
        const unsigned char* va = sV + (size_t)stage * G.stageV + baseA;
        const unsigned char* ba = sA + (size_t)stage * G.stageA + baseB;
        double2 a0[4], a1[4], b0[NTM], b1[NTM];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          a0[mt] = *reinterpret_cast<const double2*>(va + (size_t)mt * 1024 + swz0);
          a1[mt] = *reinterpret_cast<const double2*>(va + (size_t)mt * 1024 + swz1);
        }
#pragma unroll
        for (int q = 0; q < NTM; ++q)
          if (nt0 + q < G.NT) {
            b0[q] = *reinterpret_cast<const double2*>(ba + (size_t)q * 1024 + swz0);
            b1[q] = *reinterpret_cast<const double2*>(ba + (size_t)q * 1024 + swz1);
          }
        // four rounds over the warp's 4 x NTM accumulators: dependent MMAs are 4 * NTM apart
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int q = 0; q < NTM; ++q)
            if (nt0 + q < G.NT) dmma(acc[mt][q][0], acc[mt][q][1], a0[mt].x, b0[q].x);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int q = 0; q < NTM; ++q)
            if (nt0 + q < G.NT) dmma(acc[mt][q][0], acc[mt][q][1], a0[mt].y, b0[q].y);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int q = 0; q < NTM; ++q)
            if (nt0 + q < G.NT) dmma(acc[mt][q][0], acc[mt][q][1], a1[mt].x, b1[q].x);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int q = 0; q < NTM; ++q)
            if (nt0 + q < G.NT) dmma(acc[mt][q][0], acc[mt][q][1], a1[mt].y, b1[q].y);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
      }
      // ---- sum the 4 row-group pairs of every tile in hq order (fixed shape) ----
      for (int hh = 0; hh < 4; ++hh) {
        __syncthreads();
        if (hq == hh) {
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            double* urow = us + (size_t)(mt * 8 + pg) * G.J + t;
#pragma unroll
            for (int q = 0; q < NTM; ++q) {
              if (nt0 + q < G.NT) {
                double* o = urow + (nt0 + q) * 8;
                if (hh == 0) { o[0] = acc[mt][q][0]; o[4] = acc[mt][q][1]; }
                else { o[0] += acc[mt][q][0]; o[4] += acc[mt][q][1]; }
              }
            }
          }
        }
      }
    } else {
    const int box = warp >> 1;                               // row group `warp` of the chunk: box, half of its 16 rows
    const int x = ((warp & 1) << 2) | t;                     // 16-byte chunk (row pair) inside the box's lines
    const uint32_t sw = (uint32_t)((x ^ pg) << 4);
    const uint32_t offA = (uint32_t)box * 128u * (uint32_t)S::NCP + (uint32_t)pg * 128u + sw;
    const uint32_t offB = (uint32_t)box * 128u * (uint32_t)G.J + (uint32_t)pg * 128u + sw;
    double acc[MT][NTM][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTM; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    unsigned int cit = it;
    for (int c = 0; c < G.nchunks; ++c, ++cit) {
      if (!S::kDedicatedProducer) {
        if (producer && c + D < G.nchunks) issue(c + D);
        __syncwarp();
      }
      const int stage = cit % G.NS;
      const uint32_t phase = (cit / G.NS) & 1;
      mbar_wait(&full[stage], phase);
      const unsigned char* va = sV + (size_t)stage * G.stageV + offA;
      const unsigned char* ba = sA + (size_t)stage * G.stageA + offB;
      double2 a[MT], b[NTM];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = *reinterpret_cast<const double2*>(va + (size_t)mt * 1024);
#pragma unroll
      for (int nt = 0; nt < NTM; ++nt)
        if (nt < G.NT) b[nt] = *reinterpret_cast<const double2*>(ba + (size_t)nt * 1024);
      // the x halves of all tile pairs first, then the y halves: dependent MMAs on one accumulator are MT * NT apart
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTM; ++nt)
          if (nt < G.NT) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, b[nt].x);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTM; ++nt)
          if (nt < G.NT) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, b[nt].y);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
    // ---- sum the 8 row groups of every tile in warp order (fixed shape) ----
    for (int hh = 0; hh < kConsumerWarps; ++hh) {
      if (S::kDedicatedProducer) asm volatile("bar.sync 1, 256;" ::: "memory");
      else __syncthreads();
      if (warp == hh) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          double* urow = us + (size_t)(mt * 8 + pg) * G.J + t;
#pragma unroll
          for (int nt = 0; nt < NTM; ++nt) {
            if (nt < G.NT) {
              double* o = urow + nt * 8;
              if (hh == 0) { o[0] = acc[mt][nt][0]; o[4] = acc[mt][nt][1]; }
              else { o[0] += acc[mt][nt][0]; o[4] += acc[mt][nt][1]; }
            }
          }
        }
      }
    }
    }
  }
  it += (unsigned int)G.nchunks;
  __syncthreads();
}

}  // namespace tmx
}  // namespace spb
