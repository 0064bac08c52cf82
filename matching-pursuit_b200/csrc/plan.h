// plan.h -- host-side state behind an mpb200_plan_t.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "types.h"

namespace mpb {

struct Plan {
    int device = 0, sm_count = 148;
    int K = 0, A = 0, N = 0, Bmax = 0, mode = 0;
    int lo = 0, hi = 0, nloc = 0, npairs = 0;       // owned atom range
    int M = 0, blk = 0, blk_shift = 0, NB = 0;      // window FFT size, block-max granularity
    int vfull = 0, nchunks = 0;                     // full pass: blocks per window, windows per signal
    int Bcap = 0;                                   // signals resident at once (<= Bmax; map modes sub-batch)
    int NS = 0;                                     // row stride of the resident map (N, or N padded to 4 in SGRAM)
    int delta_occ = 0;                              // SGRAM: resident CTAs per SM of k_delta (persistent grid)
    int delta_occ_db = 0;                           // ... of its double-buffered two-CTA form (large resident batches)
    int M2 = 0;                                     // SGRAM: transform length of the synthesised Gram rows (>= 2A)
    bool pos_free = false;                          // SGRAM with blocks of >= 128 positions: the block/row tables carry block
                                                    // starts instead of exact positions (k_delta NOPOS); k_apply resolves
    int wcap = 0;                                   // window-spectrum slots
    int bm_cap = 0;                                 // positions refreshed by one step window (staging size)
    int cur_batch = 0;                              // batch loaded by mpb200_begin (0 = none)
    bool dict_set = false;

    float* dict = nullptr;          // (K, A) unit-normed
    C32* pairspec = nullptr;        // (npairs, M)
    C32 *tw1 = nullptr, *tw2 = nullptr;
    cpx<double> *tw1d = nullptr, *tw2d = nullptr;
    C32* pairspec2 = nullptr;       // SGRAM: (npairs, M2)
    C32* atomspec = nullptr;        // SGRAM: (K, M2) forward spectra of [0^(A-1), d_k]
    C32* tw1b = nullptr;            // BlockFft<M2> twiddles (fp32 / fp64)
    cpx<double>* tw1bd = nullptr;
    Win* win_atoms = nullptr;       // (K) windows over dictionary rows, t0 = -(A-1)
    C32* winspec = nullptr;         // (wcap, M)
    Win* win_full = nullptr;        // (Bmax * nchunks)
    Win* win_step = nullptr;        // (Bmax)
    float* bm_val = nullptr;        // (Bmax, nloc, NB)
    int* bm_pos = nullptr;
    float* row_val = nullptr;       // (Bmax, nloc)
    int* row_pos = nullptr;
    float* residual = nullptr;      // (Bmax, N)
    Best* best = nullptr;           // (Bmax)

    // local-contrast-norm selection (MPB200_OPT_LOCAL_CONTRAST_NORM, map modes): second hierarchy over the normalised map
    bool lcn = false;
    float* nbm_val = nullptr;       // (Bcap, nloc, NB)
    int* nbm_pos = nullptr;
    float* nrow_val = nullptr;      // (Bcap, nloc)
    int* nrow_pos = nullptr;

    // fused iteration loop (windowed re-correlation, latency-bound shapes): one cooperative launch per resident batch
#ifndef MPB_FUSED_DEFAULT
#define MPB_FUSED_DEFAULT 1
#endif
    bool fused_loop = MPB_FUSED_DEFAULT != 0;   // MPB200_OPT_FUSED_LOOP
    int fused_occ = 0;              // resident CTAs per SM of k_pursue_fused (0: not queried yet)
    unsigned* gbar = nullptr;       // [2] grid-barrier counters
    float* row_val2 = nullptr;      // (Bmax, nloc) second copy of the row maxima (ping-pong across iterations)
    int* row_pos2 = nullptr;

    // Gram mode
    float* gram = nullptr;          // (K, nloc, GS), GS = 2A
    float* map = nullptr;           // (Bmax, nloc, N)
    GramUpdate* upd = nullptr;      // (Bmax)
    int* trunc_count = nullptr;     // [2]
    Win* win_gram = nullptr;        // (K) windows of the Gram build
    int GS = 0;
    uint64_t gram_bytes = 0;
    unsigned iter = 0;              // iterations applied since begin (parity selects trunc_count slot)
    int refresh_every = 0;          // GRAM: full re-correlation every this many iterations (0 = never)

    // atom-sharded exchange over peer memory (NVLink): every rank owns a mailbox that all ranks write into
    int xworld = 1, xrank = 0;
    MailSlot* mail = nullptr;               // local mailbox: [2 parities][Bmax][world]
    MailSlot** peer_mail = nullptr;         // device array [world]: every rank's mailbox as seen from this device
    std::vector<void*> ipc_opened;          // peer mappings to close
    int* xerr = nullptr;                    // device flag: an exchange timed out
    unsigned xseq = 0;                      // exchanges issued so far (sequence number of the next one is xseq + 1)
    bool xconnected = false;

    // unchanged-dictionary detection (device side, no host synchronisation)
    unsigned long long* fp = nullptr;       // [0..1] fingerprint being accumulated, [2..3] previous, [4] previous valid
    int* dict_skip = nullptr;               // device flag read by the table-building kernels
    int last_normalize = -1;
    bool force_tables = false;              // MPB200_OPT_FORCE_TABLES: the next set_dictionary rebuilds whatever the fingerprint says

    // staging for the host-buffer entry point
    float* d_signal = nullptr;
    int32_t *d_atom = nullptr, *d_pos = nullptr;
    float* d_val = nullptr;
    size_t ev_cap = 0;
    void* h_stage = nullptr;

    // optional per-kernel timing (bench aid): consecutive events on the caller's stream, tagged
    // with what ran since the previous one (0 = start marker, 1 = first pass, 2 = apply, 3 = re-correlation)
    bool timing = false;
    std::vector<void*> ev_pool;         // cudaEvent_t
    std::vector<int> ev_tag;            // tags of the events recorded since the last read
    size_t ev_used = 0;

    std::vector<void*> allocs;
    uint64_t bytes = 0;
};

int fail(int code, const std::string& msg);

// Stream-ordered scratch (cudaMallocAsync) is only cheap when the device's default memory pool keeps what is freed
// into it: by default the pool hands everything back to the driver at the next synchronisation and every call pays a
// driver allocation (milliseconds).  Raises the pool's release threshold once per device.
void keep_async_pool(int device);

}  // namespace mpb
