// types.h -- small POD types shared by the kernels and the host side.
#pragma once
#include "fft_core.cuh"

namespace mpb {

using C32 = cpx<float>;

struct Win {
    int row;   // source row (signal index, or atom index for the Gram build)
    int t0;    // first sample of the window (may be negative: zero filled)
    int blk0;  // first block-max entry this window refreshes
    int nvb;   // number of blocks (of `blk` outputs) that are valid in this window
};

struct GramUpdate {   // one per signal and iteration (GRAM mode)
    float value;
    int atom;      // global atom index
    int position;
    int valid;     // 1: apply the Gram update; 0: this signal takes the FFT route
};

struct Best {  // == mpb200_best
    float value;
    int atom;
    int position;
    int pad;
};

}  // namespace mpb
