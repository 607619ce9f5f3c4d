// eqf_gemm_tf32x3.cu - hand-written tcgen05 GEMM for the per-degree channel-mixing linears (sm_100a).
//
//   C[M, N] = A[M, K] * Bt[N, K]^T        fp32 in HBM, fp32-level accuracy, M = edges x (2l+1) rows (tall), K, N <= ~1000
//
// The linears after each depth-wise tensor product (LinearRS, nets/tensor_product_rescale.py:165-174; in the reference
// an e3nn 'uvw' einsum -> cuBLAS SGEMM) are skinny: tens of thousands of rows, K and N of a few hundred.  The CUTLASS
// 9xBF16 collective (eqf_gemm.cu) keeps the tensor pipe 33 % busy on them: every value is split three ways by a
// transform warp-group and the TMEM accumulator is promoted to registers every other k-block.  This kernel uses the
// 3xTF32 scheme instead:
//   a = a_hi + a_lo,  a_hi = the top 19 bits of a (what a kind::tf32 MMA reads from a raw fp32 operand),
//   a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi          (the dropped a_lo*b_lo term is 2^-22 relative)
// so the *raw* TMA-written fp32 tile is the hi operand as is, the only transform is lo = x - trunc19(x) on the A tile
// (the weights' hi / lo planes are split once per call by a tiny kernel), and all three products accumulate in one
// TMEM accumulator over the whole K loop - no promotion.
//
// One CTA per SM, persistent over 128 x BN output tiles, warp-specialised:
//   warps 0-3  epilogue   (tcgen05.ld TMEM -> registers -> st.global; warp w owns TMEM lanes 32w..32w+31)
//   warp  4    TMA producer (one lane): A raw tile, B hi tile, B lo tile per 32-wide k-tile (128-byte rows, SWIZZLE_128B)
//   warp  5    MMA issuer (one lane): 4 k-blocks x 3 tcgen05.mma.kind::tf32 per k-tile; owns the TMEM allocation
//   warps 6-9  transform: A_lo tile from the A raw tile (in shared memory, same swizzled positions)
// Pipelines: smem ring (full / lo_ready / empty mbarriers) and a double-buffered TMEM accumulator (tmem_full / tmem_empty),
// so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>

#include "eqf_common.cuh"

namespace eqf {
namespace tf32x3 {

constexpr int BM = 128;          // rows per tile (UMMA M)
constexpr int BK = 32;           // fp32 per k-tile row = 128 bytes = one SWIZZLE_128B atom
constexpr int UMMA_K = 8;        // tf32 MMA depth
constexpr int kThreads = 320;
constexpr int kEpilogueWarps = 4, kProducerWarp = 4, kMmaWarp = 5, kTransformWarp0 = 6, kTransformWarps = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives on `bar` once every MMA issued so far has finished
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile in shared memory: rows of 128 bytes, 8-row groups 1024 bytes apart, SWIZZLE_128B
// (cute UMMA::SmemDescriptor: start>>4 | LBO(=1)<<16 | SBO(=64)<<32 | version(=1)<<46 | layout SWIZZLE_128B(=2)<<61)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32, fp32 accumulate, K-major A and B, M = 128, N = n  (cute UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t instr_desc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

struct Params {
  float* C;
  long long M, N, K, ldc;
  int n_tile;        // columns per output tile (multiple of 16, <= BN)
  int n_blocks;      // ceil(N / n_tile)
  long long m_blocks;
};

template <int BN>
struct Smem {
  static constexpr int kStages = (BN >= 256) ? 2 : (BN >= 128 ? 3 : 4);
  static constexpr int kABytes = BM * BK * 4;      // 16 KB
  static constexpr int kBBytes = BN * BK * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024 /* alignment slack */;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                   const __grid_constant__ CUtensorMap map_blo, Params p) {
  using S = Smem<BN>;
  constexpr int kStages = S::kStages;
  constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* full = bars;                       // [kStages] TMA landed
  uint64_t* lo_ready = bars + kStages;         // [kStages] A_lo written
  uint64_t* empty = bars + 2 * kStages;        // [kStages] MMAs reading the stage have finished
  uint64_t* tmem_full = bars + 3 * kStages;    // [2]
  uint64_t* tmem_empty = bars + 3 * kStages + 2;  // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long k_tiles = (p.K + BK - 1) / BK;
  const long long n_tiles_total = p.m_blocks * p.n_blocks;

  if (warp == kProducerWarp && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_bhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_blo)) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&lo_ready[s], kTransformWarps);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kEpilogueWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {   // whole warp: TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == kProducerWarp) {
    // ===================================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      const uint32_t tx = (uint32_t)(S::kABytes + 2 * p.n_tile * BK * 4);
      for (long long tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
        const long long mb = tile / p.n_blocks;
        const int nb = (int)(tile % p.n_blocks);
        for (long long kt = 0; kt < k_tiles; ++kt, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = stage_base + (size_t)s * S::kStageBytes;
          mbar_expect_tx(&full[s], tx);
          tma_load_2d(st, &map_a, (int)(kt * BK), (int)(mb * BM), &full[s]);
          tma_load_2d(st + 2 * S::kABytes, &map_bhi, (int)(kt * BK), nb * p.n_tile, &full[s]);
          tma_load_2d(st + 2 * S::kABytes + S::kBBytes, &map_blo, (int)(kt * BK), nb * p.n_tile, &full[s]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = instr_desc(p.n_tile);
      uint32_t it = 0, acc_it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++acc_it) {
        const int a = acc_it & 1;
        const uint32_t aph = (acc_it >> 1) & 1;
        mbar_wait(&tmem_empty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
        for (long long kt = 0; kt < k_tiles; ++kt, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&full[s], ph);
          mbar_wait(&lo_ready[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(stage_base + (size_t)s * S::kStageBytes);
          const uint64_t a_hi = smem_desc(st), a_lo = smem_desc(st + S::kABytes);
          const uint64_t b_hi = smem_desc(st + 2 * S::kABytes), b_lo = smem_desc(st + 2 * S::kABytes + S::kBBytes);
#pragma unroll
          for (int kb = 0; kb < BK / UMMA_K; ++kb) {
            const uint64_t adv = (uint64_t)((kb * UMMA_K * 4) >> 4);     // 32 bytes per k-block inside the 128-byte row
            umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (kt > 0 || kb > 0) ? 1u : 0u);
            umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
          }
          umma_commit(&empty[s]);                    // frees the smem stage once these MMAs have read it
        }
        umma_commit(&tmem_full[a]);                  // accumulator complete
      }
    }
  } else if (warp >= kTransformWarp0) {
    // ===================================================================================== transform: A_lo = A - trunc19(A)
    const int t = threadIdx.x - kTransformWarp0 * 32;   // 0..127
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
      for (long long kt = 0; kt < k_tiles; ++kt, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full[s], ph);
        const float4* raw = reinterpret_cast<const float4*>(stage_base + (size_t)s * S::kStageBytes);
        float4* lo = reinterpret_cast<float4*>(stage_base + (size_t)s * S::kStageBytes + S::kABytes);
#pragma unroll
        for (int i = 0; i < (BM * BK / 4) / (kTransformWarps * 32); ++i) {
          const int idx = i * (kTransformWarps * 32) + t;
          const float4 v = raw[idx];
          float4 r;
          r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
          r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
          r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          lo[idx] = r;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&lo_ready[s]);
      }
    }
  } else {
    // ===================================================================================== epilogue (warps 0-3)
    uint32_t acc_it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++acc_it) {
      const long long mb = tile / p.n_blocks;
      const int nb = (int)(tile % p.n_blocks);
      const int a = acc_it & 1;
      const uint32_t aph = (acc_it >> 1) & 1;
      mbar_wait(&tmem_full[a], aph);
      tc_fence_after();
      const long long row = mb * BM + warp * 32 + lane;
      const long long col0 = (long long)nb * p.n_tile;
      float* crow = p.C + row * p.ldc + col0;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * BN);
      for (int c = 0; c < p.n_tile; c += 16) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr + (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < p.M) {
          if (col0 + c + 16 <= p.N) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(crow + c + 4 * q) =
                  make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                              __uint_as_float(v[4 * q + 3]));
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (col0 + c + q < p.N) crow[c + q] = __uint_as_float(v[q]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[a]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// hi / lo planes of the (small) weight operand: hi = top 19 bits, lo = w - hi
__global__ void split_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = w[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled encode_fn() {
  static EncodeTiled fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiled>(f);
  }();
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row stride ld (elements), box = [box_rows, 32 columns], 128-byte swizzle, zero OOB fill
static int make_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiled enc = encode_fn();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EQF_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")"); return EQF_ERR_CUDA; }
  return EQF_OK;
}

template <int BN>
static int launch(const CUtensorMap& ma, const CUtensorMap& mh, const CUtensorMap& ml, const Params& p, cudaStream_t s) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tf32x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::kTotal);
  });
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "gemm_tf32x3 smem attribute");
  int sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = p.m_blocks * p.n_blocks;
  const int grid = (int)(tiles < sms ? tiles : sms);
  gemm_tf32x3_kernel<BN><<<grid, kThreads, Smem<BN>::kTotal, s>>>(ma, mh, ml, p);
  return check_cuda(cudaGetLastError(), "gemm_tf32x3_kernel launch");
}

}  // namespace tf32x3
}  // namespace eqf

using namespace eqf;

// C[M, N] = A[M, K] (row-major, lda) x Bt[N, K]^T (row-major, ldb), 3xTF32 on tcgen05.  `split` is device scratch of
// 2 * N * K floats for the hi / lo planes of Bt.  All pointers 16-byte aligned, K, lda, ldb, ldc multiples of 4.
extern "C" int eqf_gemm_tf32x3(const float* A, const float* Bt, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                               int64_t ldb, int64_t ldc, float* split, void* stream) {
  using namespace eqf::tf32x3;
  if (M <= 0 || N <= 0) return EQF_OK;
  if (!A || !Bt || !C || !split) { set_error("eqf_gemm_tf32x3: null pointer"); return EQF_ERR_INVALID; }
  if (K <= 0) { set_error("eqf_gemm_tf32x3: K must be positive"); return EQF_ERR_INVALID; }
  if ((((uintptr_t)A | (uintptr_t)Bt | (uintptr_t)C | (uintptr_t)split) & 15) || ((K | lda | ldb | ldc) & 3) || lda < K ||
      ldb < K || ldc < N) {
    set_error("eqf_gemm_tf32x3: operands must be 16-byte aligned, K and leading dimensions multiples of 4");
    return EQF_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  // hi / lo planes of the weights, packed [N, K]
  float* hi = split;
  float* lo = split + N * K;
  if (ldb == K) {
    split_kernel<<<(unsigned)((N * K + 255) / 256 < 1184 ? (N * K + 255) / 256 : 1184), 256, 0, s>>>(Bt, hi, lo, N * K);
  } else {
    set_error("eqf_gemm_tf32x3: Bt must be packed (ldb == K)");
    return EQF_ERR_UNSUPPORTED;
  }
  int rc = check_cuda(cudaGetLastError(), "split_kernel launch");
  if (rc != EQF_OK) return rc;
  // tile the columns: at most 256 per tile, multiples of 16, as even as possible
  const int n_blocks = (int)((N + 255) / 256);
  int n_tile = (int)((N + n_blocks - 1) / n_blocks);
  n_tile = (n_tile + 15) & ~15;
  Params p;
  p.C = C; p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.n_tile = n_tile; p.n_blocks = (int)((N + n_tile - 1) / n_tile);
  p.m_blocks = (M + BM - 1) / BM;
  CUtensorMap ma, mh, ml;
  if ((rc = make_map(&ma, A, M, K, lda, BM)) != EQF_OK) return rc;
  if ((rc = make_map(&mh, hi, N, K, K, n_tile)) != EQF_OK) return rc;
  if ((rc = make_map(&ml, lo, N, K, K, n_tile)) != EQF_OK) return rc;
  if (n_tile <= 64) return launch<64>(ma, mh, ml, p, s);
  if (n_tile <= 128) return launch<128>(ma, mh, ml, p, s);
  return launch<256>(ma, mh, ml, p, s);
}
