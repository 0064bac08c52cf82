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

// One rank's candidate for one signal in another rank's mailbox (atom-sharded exchange over peer memory).
// Every 8-byte word carries (payload, sequence number), so a word is either entirely old or entirely new
// (8-byte stores are single-copy atomic): no fence and no separate flag, as in NCCL's LL protocol.
struct MailSlot {
    unsigned long long w[4];   // [0] value bits | seq << 32, [1] atom | seq << 32, [2] position | seq << 32, [3] unused
};

struct Best {  // == mpb200_best
    float value;
    int atom;
    int position;
    int pad;
};

}  // namespace mpb
