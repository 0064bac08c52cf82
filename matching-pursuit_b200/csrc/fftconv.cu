// fftconv.cu -- two-operand zero-padded FFT convolution / correlation
// (modules/fft.py:23-35, modules/transfer.py:548-569).
#include "../../include/mpb200.h"
#include "plan.h"

using namespace mpb;

extern "C" int mpb200_fft_convolve(const float* a, int rows_a, const float* b, int rows_b, int n, int conjugate_b,
                                   float* out, void* stream) {
    (void)a; (void)rows_a; (void)b; (void)rows_b; (void)n; (void)conjugate_b; (void)out; (void)stream;
    return fail(MPB200_EINVAL, "mpb200_fft_convolve: not built into this library version");
}
