// eqf_gemm.cu - fp32-accurate tensor-core GEMM for the per-degree channel-mixing linears (sm_100a, tcgen05).
//
// The linears that follow each depth-wise tensor product (SeparableFCTP.lin / sep_alpha / proj / merge / FFN:
// LinearRS, nets/tensor_product_rescale.py:165-174 -> e3nn 'uvw' einsum -> cuBLAS SGEMM in the reference) are plain
// row-major GEMMs on the planar buffers: [rows*(2l+1), K_l] x [K_l, C_l].  They hold the FLOP majority of the layer
//  and fp32 parity (1e-4) rules out single-pass TF32, so this file instantiates
// CUTLASS's sm_100 "FastFP32" collective (vendored CUTLASS 4.5 headers): operands stay fp32 in HBM, TMA stages them
// into shared memory, a transform warp-group splits every value into three bf16 terms, tcgen05.mma accumulates the
// five significant cross products in TMEM (fp32 accuracy), and the epilogue reads TMEM back with tcgen05.ld.
// This is library code in the sense of the task statement (like calling cuBLAS) - it replaces cuBLAS's SIMT SGEMM,
// which ran at ~21 TFLOP/s on these skinny shapes.
//
// Three operand layouts cover forward, data gradient and weight gradient of Y = X W (all tensors row-major):
//   NN: C[M,N] = A[M,K]  B[K,N]      TN-like dgrad: C[M,K'] = A[M,N'] B[K',N']^T     wgrad: C[K,N] = A[M,K]^T B[M,N]
#include <cuda_runtime.h>

#include "cutlass/cutlass.h"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/util/packed_stride.hpp"
#include "cute/tensor.hpp"

#include <cstdlib>
#include <string>

#include "../../include/eqf_b200.h"

// self-contained (built into its own libeqf_gemm.so so that the CUTLASS instantiations are compiled once)
namespace eqf {

using namespace cute;

static thread_local std::string g_gemm_error;
static void set_error(const std::string& msg) { g_gemm_error = msg; }
static int check_cuda(cudaError_t err, const char* what) {
  if (err == cudaSuccess) return EQF_OK;
  g_gemm_error = std::string(what) + ": " + cudaGetErrorString(err);
  return EQF_ERR_CUDA;
}

// The stock builder fixes NumBandsToCompute = 5 (all nine bf16 cross products) and AccPromotionInterval = 1.
// Bands 4 and 5 (A1*B2 + A2*B1, A2*B2) carry terms below 2^-24 of the product - under the fp32 rounding of the
// result - so the policy is rebound here with a compile-time band count (3 keeps six of the nine MMAs).
// Measured on the layer shapes (profiles/r1_gemm_variants.md): the band count barely matters (the kernel is not
// MMA-bound), promoting the TMEM accumulator into registers every 2 k-blocks instead of every one is worth 20 %;
// rel. error vs fp64 stays 2.7e-7 (cuBLAS SGEMM: 8e-7).  K tile 64 / interval 4 does not fit shared memory.
#ifndef EQF_GEMM_BANDS
#define EQF_GEMM_BANDS 3
#endif
#ifndef EQF_GEMM_PROMO
#define EQF_GEMM_PROMO 2
#endif
#ifndef EQF_GEMM_TILEK
#define EQF_GEMM_TILEK 32
#endif

template <class Op, int Bands, int Promo>
struct RebindBands { using type = Op; };

template <int L2T, int T2M, int Sch, int Acc, int NB, int SF, int API, class CS, class ACA, class Arch, int Bands,
          int Promo, class... Rest>
struct RebindBands<cutlass::gemm::collective::CollectiveMma<
                       cutlass::gemm::MainloopSm100TmaUmmaWarpSpecializedFastF32<L2T, T2M, Sch, Acc, NB, SF, API, CS, ACA, Arch>,
                       Rest...>,
                   Bands, Promo> {
  using type = cutlass::gemm::collective::CollectiveMma<
      cutlass::gemm::MainloopSm100TmaUmmaWarpSpecializedFastF32<L2T, T2M, Sch, Acc, Bands, SF, Promo, CS, ACA, Arch>, Rest...>;
};

template <class LayoutA, class LayoutB, int TileN, int TileK = EQF_GEMM_TILEK, class Scheduler = void, bool TwoSm = false>
struct FastF32Gemm {
  using ElementA = float;
  using ElementB = float;
  using ElementC = float;
  using ElementAcc = float;
  using LayoutC = cutlass::layout::RowMajor;
  static constexpr int Align = 4;  // 128-bit
  // 1-SM: 128 x N MMA atoms, no cluster; 2-SM: cta_group::2 pairs (256 x N per pair), cluster 2x1
  using MmaTile = Shape<Int<TwoSm ? 256 : 128>, Int<TileN>, Int<TileK>>;
  using Cluster = Shape<Int<TwoSm ? 2 : 1>, _1, _1>;
  using Schedule = cute::conditional_t<TwoSm, cutlass::gemm::KernelTmaWarpSpecialized2SmFastFP32Sm100,
                                       cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32Sm100>;

  using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTile, Cluster,
      cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc, ElementC, LayoutC, Align, ElementC,
      LayoutC, Align, cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;

  using StockMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutA, Align, ElementB, LayoutB, Align,
      ElementAcc, MmaTile, Cluster,
      cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
      Schedule>::CollectiveOp;
  using CollectiveMainloop = typename RebindBands<StockMainloop, EQF_GEMM_BANDS, EQF_GEMM_PROMO>::type;
  static_assert(CollectiveMainloop::DispatchPolicy::NumBandsToCompute == EQF_GEMM_BANDS, "band rebinding did not apply");

  using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, Scheduler>;
  using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;

  using StrideA = typename Gemm::GemmKernel::StrideA;
  using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC;
  using StrideD = typename Gemm::GemmKernel::StrideD;

  // batch > 1: `batch` independent products with operand strides (batch_a, batch_b, batch_c) elements apart -
  // used to split the row reduction of the weight gradient into slices that fill the machine.
  static int run(const float* A, const float* B, float* C, int M, int N, int K, long long lda, long long ldb,
                 long long ldc, float beta, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                 int batch = 1, long long batch_a = 0, long long batch_b = 0, long long batch_c = 0) {
    // leading dimensions: A row-major -> lda = elements between rows of A[M,K]; column-major -> between columns
    StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, make_shape(M, K, batch));
    StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, make_shape(N, K, batch));
    StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, make_shape(M, N, batch));
    set_ld(sa, lda);
    set_ld(sb, ldb);
    set_ld(sc, ldc);
    if (batch > 1) { get<2>(sa) = batch_a; get<2>(sb) = batch_b; get<2>(sc) = batch_c; }
    typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm,
                                  {M, N, K, batch},
                                  {A, sa, B, sb},
                                  {{1.0f, beta}, C, sc, C, sc}};
    Gemm gemm;
    if (gemm.can_implement(args) != cutlass::Status::kSuccess) {
      set_error("fast-fp32 GEMM: shape/alignment not supported (dims and leading dims must be multiples of 4)");
      return EQF_ERR_UNSUPPORTED;
    }
    size_t need = Gemm::get_workspace_size(args);
    if (need > workspace_bytes) { set_error("fast-fp32 GEMM: workspace too small"); return EQF_ERR_INVALID; }
    if (gemm.initialize(args, workspace, stream) != cutlass::Status::kSuccess) {
      set_error("fast-fp32 GEMM: initialize failed"); return EQF_ERR_CUDA;
    }
    if (gemm.run(stream) != cutlass::Status::kSuccess) { set_error("fast-fp32 GEMM: launch failed"); return EQF_ERR_CUDA; }
    return check_cuda(cudaGetLastError(), "fast-fp32 GEMM launch");
  }

  // packed strides are (ld, 1, batch) or (1, ld, batch): overwrite whichever mode is the non-unit one
  template <class Stride>
  static void set_ld(Stride& s, long long ld) {
    if constexpr (cute::is_static_v<decltype(get<0>(s))>) {
      get<1>(s) = ld;
    } else {
      get<0>(s) = ld;
    }
  }
};

}  // namespace eqf

using namespace eqf;
using Row = cutlass::layout::RowMajor;
using Col = cutlass::layout::ColumnMajor;

// mode 0: C[M,N] = A[M,K] (row-major, lda) x B[K,N] (row-major, ldb)
// mode 1: C[M,N] = A[M,K] (row-major, lda) x B^T where B is [N,K] row-major (ldb)          (data gradient)
// mode 2: C[M,N] = A^T x B where A is [K,M] row-major (lda) and B is [K,N] row-major (ldb)   (weight gradient)
// beta = 0 overwrites C, beta = 1 accumulates.  Workspace: device scratch of eqf_gemm_workspace_bytes().
extern "C" int eqf_gemm_f32(int mode, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                            int64_t lda, int64_t ldb, int64_t ldc, float beta, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  if (A == nullptr || B == nullptr || C == nullptr) { set_error("eqf_gemm_f32: null pointer"); return EQF_ERR_INVALID; }
  if (M <= 0 || N <= 0 || K <= 0) return EQF_OK;
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15 || (lda | ldb | ldc) & 3) {
    set_error("eqf_gemm_f32: operands must be 16-byte aligned with leading dimensions that are multiples of 4 floats");
    return EQF_ERR_INVALID;
  }
  if (M > 0x7fffffffLL || N > 0x7fffffffLL || K > 0x7fffffffLL) { set_error("eqf_gemm_f32: dimension too large"); return EQF_ERR_UNSUPPORTED; }
  cudaStream_t s = (cudaStream_t)stream;
  const int m = (int)M, n = (int)N, k = (int)K;
  const bool wide = n > 64;
  // cta_group::2 tiles measured 3-5 % faster than 1-SM tiles on every layer shape (profiles/r1_gemm_microbench*.jsonl)
  static const bool two_sm = [] { const char* e = std::getenv("EQF_GEMM_2SM"); return e == nullptr || e[0] != '0'; }();
  if (two_sm && mode != 2) {
    if (mode == 0)
      return wide ? FastF32Gemm<Row, Row, 128, EQF_GEMM_TILEK, void, true>::run(A, B, C, m, n, k, lda, ldb, ldc, beta, workspace, workspace_bytes, s)
                  : FastF32Gemm<Row, Row, 64, EQF_GEMM_TILEK, void, true>::run(A, B, C, m, n, k, lda, ldb, ldc, beta, workspace, workspace_bytes, s);
    return wide ? FastF32Gemm<Row, Col, 128, EQF_GEMM_TILEK, void, true>::run(A, B, C, m, n, k, lda, ldb, ldc, beta, workspace, workspace_bytes, s)
                : FastF32Gemm<Row, Col, 64, EQF_GEMM_TILEK, void, true>::run(A, B, C, m, n, k, lda, ldb, ldc, beta, workspace, workspace_bytes, s);
  }
  switch (mode) {
    case 0:
      return wide ? FastF32Gemm<Row, Row, 128>::run(A, B, C, m, n, k, lda, ldb, ldc, beta, workspace, workspace_bytes, s)
                  : FastF32Gemm<Row, Row, 64>::run(A, B, C, m, n, k, lda, ldb, ldc, beta, workspace, workspace_bytes, s);
    case 1:
      return wide ? FastF32Gemm<Row, Col, 128>::run(A, B, C, m, n, k, lda, ldb, ldc, beta, workspace, workspace_bytes, s)
                  : FastF32Gemm<Row, Col, 64>::run(A, B, C, m, n, k, lda, ldb, ldc, beta, workspace, workspace_bytes, s);
    case 2:
      return wide ? FastF32Gemm<Col, Row, 128>::run(A, B, C, m, n, k, lda, ldb, ldc, beta, workspace, workspace_bytes, s)
                  : FastF32Gemm<Col, Row, 64>::run(A, B, C, m, n, k, lda, ldb, ldc, beta, workspace, workspace_bytes, s);
    default:
      set_error("eqf_gemm_f32: mode must be 0, 1 or 2");
      return EQF_ERR_INVALID;
  }
}

extern "C" int64_t eqf_gemm_workspace_bytes(void) { return 64 << 20; }

extern "C" const char* eqf_gemm_last_error(void) { return g_gemm_error.c_str(); }


// Weight gradient with the row reduction split into `slices` equal chunks of `chunk` rows (slices * chunk <= K rows):
// part[s][M, N] = A[s*chunk : (s+1)*chunk, :M]^T  B[s*chunk : (s+1)*chunk, :N]   for s < slices  (one batched launch of
// the tcgen05 fast-fp32 kernel - the batch dimension supplies the parallelism the tiny [M, N] output lacks).
// The caller reduces `part` over s (and adds the tail rows, if any, with one more mode-2 call).
extern "C" int eqf_gemm_f32_wgrad_sliced(const float* A, const float* B, float* part, int64_t M, int64_t N, int64_t chunk,
                                         int64_t slices, int64_t lda, int64_t ldb, void* workspace,
                                         int64_t workspace_bytes, void* stream) {
  if (A == nullptr || B == nullptr || part == nullptr) { set_error("eqf_gemm_f32_wgrad_sliced: null pointer"); return EQF_ERR_INVALID; }
  if (M <= 0 || N <= 0 || chunk <= 0 || slices <= 0) return EQF_OK;
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)part) & 15 || (lda | ldb | N) & 3 || (chunk & 3)) {
    set_error("eqf_gemm_f32_wgrad_sliced: operands must be 16-byte aligned (dims / chunk multiples of 4)");
    return EQF_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int m = (int)M, n = (int)N, k = (int)chunk, L = (int)slices;
  if (n > 64)
    return FastF32Gemm<Col, Row, 128>::run(A, B, part, m, n, k, lda, ldb, N, 0.f, workspace, workspace_bytes, s, L,
                                           chunk * lda, chunk * ldb, M * N);
  return FastF32Gemm<Col, Row, 64>::run(A, B, part, m, n, k, lda, ldb, N, 0.f, workspace, workspace_bytes, s, L,
                                        chunk * lda, chunk * ldb, M * N);
}


extern "C" int eqf_gemm_config(int* bands, int* promo, int* tile_k) {
  if (bands) *bands = EQF_GEMM_BANDS;
  if (promo) *promo = EQF_GEMM_PROMO;
  if (tile_k) *tile_k = EQF_GEMM_TILEK;
  return EQF_OK;
}
