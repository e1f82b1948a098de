// Cross-check of the hand-packed tcgen05 descriptors of the EXPERIMENTAL tensor-core encoder against CuTe's own encoders
// (vendored CUTLASS headers; host-side run, no GPU):
//   nvcc -std=c++17 -I<cutlass/include> -arch=sm_100a -o cute_desc_check tools/cute_desc_check.cu && ./cute_desc_check
// Prints, one per line:  idesc <M> <N> <value>   and   operand <rows> <lbo_16B> <sbo_16B> <layout_type> <version>
// for the instruction descriptors (TF32 x TF32 -> F32, K-major A and B) and for the K-major no-swizzle operand layout
//   element (row, k) -> (k / 4) * rows * 4 + row * 4 + (k % 4)   [floats]      (one K-step: k < 8)
// that m6a_encoder_tc.cu uses for X (128 rows), W1 (160 rows) and W2 (32 rows).  cute::UMMA::make_umma_desc statically
// asserts that the layout is a canonical UMMA K-major layout and derives LBO / SBO from it.
// tests/test_encoder_tc.py compares the output with m6a_tc_geometry().
#include <cstdio>

#include <cute/tensor.hpp>
#include <cute/arch/mma_sm100_desc.hpp>
#include <cute/atom/mma_traits_sm100.hpp>

using namespace cute;

template <int ROWS>
void operand() {
  auto layout = make_layout(make_shape(Int<ROWS>{}, make_shape(Int<4>{}, Int<2>{})),
                            make_stride(Int<4>{}, make_stride(Int<1>{}, Int<ROWS * 4>{})));
  alignas(128) static tfloat32_t buf[ROWS * 8];
  auto t = make_tensor(make_smem_ptr(buf), layout);
  auto d = UMMA::make_umma_desc<UMMA::Major::K>(t);     // start address is meaningless on the host
  printf("operand %d %u %u %u %u\n", ROWS, (unsigned)d.leading_byte_offset_, (unsigned)d.stride_byte_offset_,
         (unsigned)d.layout_type_, (unsigned)d.version_);
}

int main() {
  auto d1 = UMMA::make_instr_desc<tfloat32_t, tfloat32_t, float, 128, 160, UMMA::Major::K, UMMA::Major::K>();
  auto d2 = UMMA::make_instr_desc<tfloat32_t, tfloat32_t, float, 128, 32, UMMA::Major::K, UMMA::Major::K>();
  printf("idesc 128 160 %u\nidesc 128 32 %u\n", d1.desc_, d2.desc_);
  operand<128>();
  operand<160>();
  operand<32>();
  return 0;
}
