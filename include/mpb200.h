/* mpb200.h -- C ABI of the B200-native greedy matching-pursuit engine.
 *
 * The reference (JohnVinyard/matching-pursuit) is pure Python/PyTorch and has
 * no FFI; the boundary it offers is a set of Python callables (SURVEY.md
 * section 8b).  This header is the C-level seam those callables are re-hosted
 * on: every entry point names the reference interface it replaces (paths are
 * relative to the reference tree).  Signatures carry plain pointers and sizes
 * only -- no torch types.  Unless a function says "host", every pointer is a
 * DEVICE pointer to contiguous memory, work is enqueued on `stream` (a
 * cudaStream_t passed as void*; NULL = legacy default stream) and the call
 * returns without synchronising.  All arithmetic is fp32; indices are int32
 * (positions < 2^31, atoms < 2^31).
 *
 * Return value: 0 on success, a negative MPB200_E* code otherwise;
 * mpb200_last_error() gives the message for the calling thread.
 */
#ifndef MPB200_H
#define MPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPB200_VERSION 130   /* 1.3: option FUSED_LOOP (one cooperative launch per pursuit in the windowed re-correlation mode) */

#define MPB200_OK 0
#define MPB200_EINVAL (-1)   /* bad argument / unsupported shape */
#define MPB200_ECUDA (-2)    /* CUDA runtime error */
#define MPB200_ENOMEM (-3)   /* device allocation failed */
#define MPB200_ESTATE (-4)   /* call sequence error (no dictionary set, ...) */

/* How the correlation map is kept up to date after the first full pass. */
#define MPB200_MODE_AUTO 0        /* GRAM when the table fits and is amortised, else SGRAM when the map of a useful
                                     sub-batch fits, else RECORRELATE */
#define MPB200_MODE_RECORRELATE 1 /* re-correlate the +-A window of each winner by FFT; only block maxima are resident */
#define MPB200_MODE_GRAM 2        /* resident map, updated from a precomputed atom cross-correlation table */
#define MPB200_MODE_FULL 3        /* recompute the whole map every step (the reference's schedule) */
#define MPB200_MODE_SGRAM 4       /* resident map, updated from Gram rows that are SYNTHESISED per step from cached
                                     atom spectra (one inverse FFT of >= 2A points per atom pair); the K^2(2A-1)
                                     table is never stored.  Batches larger than the resident capacity are
                                     processed in sub-batches. */

typedef struct mpb200_plan* mpb200_plan_t;

typedef struct mpb200_plan_info {
    int32_t n_atoms, atom_size, n_samples, max_batch;
    int32_t mode;            /* resolved mode (never AUTO) */
    int32_t fft_size;        /* M: window FFT length */
    int32_t block;           /* positions per block-max entry */
    int32_t n_blocks;        /* ceil(n_samples / block) */
    int32_t atom_lo, atom_hi;/* atoms owned by this plan (atom sharding) */
    int32_t resident_batch;  /* signals processed at once (sub-batch size of the map modes) */
    int32_t fft_size2;       /* SGRAM: transform length of the synthesised Gram rows */
    uint64_t device_bytes;   /* device memory owned by the plan */
    uint64_t gram_bytes;     /* of which: Gram table */
} mpb200_plan_info;

/* One selection record, 16 bytes; what atom-sharded ranks exchange per step. */
typedef struct mpb200_best {
    float value;
    int32_t atom;     /* global atom index */
    int32_t position;
    int32_t pad;
} mpb200_best;

int mpb200_version(void);
const char* mpb200_last_error(void);
/* Number of CUDA kernels this library has launched in this process (monotonic;
 * benchmark bookkeeping, no reference counterpart). */
unsigned long long mpb200_launch_count(void);

/* Plan: sizes workspaces for signals of n_samples, batches up to max_batch and
 * a dictionary of n_atoms x atom_size of which this plan owns atoms
 * [atom_lo, atom_hi) (pass 0, n_atoms for no sharding).  gram_budget_bytes
 * bounds the Gram table in AUTO mode (0 = default 40% of free memory); in
 * SGRAM mode it bounds the resident correlation map instead (0 = 85% of free
 * memory), which sets plan_info.resident_batch: larger batches are processed
 * in balanced sub-batches of at most that many signals.
 * No reference counterpart: the reference re-derives everything per call
 * (modules/matchingpursuit.py:254-259). */
int mpb200_plan_create(mpb200_plan_t* plan, int n_atoms, int atom_size, int n_samples, int max_batch,
                       int mode, int atom_lo, int atom_hi, uint64_t gram_budget_bytes);
int mpb200_plan_destroy(mpb200_plan_t plan);
int mpb200_plan_info_get(mpb200_plan_t plan, mpb200_plan_info* info);
/* Options.  MPB200_OPT_REFRESH_EVERY (GRAM and SGRAM modes): re-correlate the whole map from
 * the residual every `value` iterations to bound the drift of the incremental
 * fp32 updates (0 = never, the default). */
#define MPB200_OPT_REFRESH_EVERY 1
/* MPB200_OPT_FORCE_TABLES (value != 0): the next mpb200_plan_set_dictionary* rebuilds every derived table even if
 * the dictionary's fingerprint equals the previous one (see below). */
#define MPB200_OPT_FORCE_TABLES 2
/* MPB200_OPT_MAX_STEPS: size the device staging of mpb200_sparse_code_host for `value` iterations per signal now
 * (plans are created with room for MPB200_DEFAULT_MAX_STEPS; a host call that needs more re-sizes on the spot,
 * the only allocation the library makes after plan creation). */
#define MPB200_OPT_MAX_STEPS 3
#define MPB200_DEFAULT_MAX_STEPS 1024
/* MPB200_OPT_POSITION_FREE (SGRAM mode, blocks of >= 128 positions): 1 = the block/row maxima tables carry block
 * starts instead of exact positions and the kernel that applies a winner resolves its exact position from the
 * resident map; 0 = exact positions everywhere.  Chosen automatically at plan creation (on when the refresh kernel
 * dominates: >= 1024 (atom pair, signal) work items per iteration); results are identical either way.  Takes
 * effect from the next mpb200_begin / mpb200_sparse_code. */
#define MPB200_OPT_POSITION_FREE 4
/* MPB200_OPT_LOCAL_CONTRAST_NORM (GRAM / SGRAM mode, un-sharded plans): 1 = the selection of mpb200_sparse_code runs
 * on fm - avg_pool2d(fm, 9x9, stride 1, zero padding 4) over the (atom, time) plane and the reported value is the
 * raw map value at the winner (modules/matchingpursuit.py:286-296), INCREMENTALLY: after every step only the
 * normalised values in the winner's +-A window (+-4 columns) are recomputed from the resident map and a second
 * block/row-max hierarchy is refreshed.  Allocates that hierarchy on first use. */
#define MPB200_OPT_LOCAL_CONTRAST_NORM 5
/* MPB200_OPT_FUSED_LOOP (windowed re-correlation mode, un-sharded plans): 1 (default) = mpb200_sparse_code runs the
 * whole iteration loop of a resident batch as ONE cooperative launch (pair spectra in registers, one grid barrier
 * per iteration) whenever one CTA per (atom-pair group, signal) fits on the device at once -- the latency-bound
 * shapes, e.g. BASELINE configs[0]; 0 = always the stream-ordered loop of two launches per iteration.  Both give
 * the same events bit for bit (reference loop: modules/matchingpursuit.py:298-328). */
#define MPB200_OPT_FUSED_LOOP 6
int mpb200_plan_set_option(mpb200_plan_t plan, int option, long long value);

/* Per-kernel device timing of the pursuit loop (bench / profiling aid, no
 * reference counterpart).  While enabled, mpb200_sparse_code records a CUDA
 * event on `stream` after the first full pass and after every apply and
 * re-correlation launch.  mpb200_plan_timing_read waits for the last event and
 * returns, for tag 1 = first pass, 2 = select+subtract+window FFT ("apply"),
 * 3 = window re-correlation + block/row maxima, 4 = Gram-table map update
 * (GRAM mode) / synthesised Gram-row map update (SGRAM mode), the accumulated milliseconds and number of intervals since the
 * previous read (arrays of 5; index 0 unused). */
int mpb200_plan_timing_enable(mpb200_plan_t plan, int enable);
int mpb200_plan_timing_read(mpb200_plan_t plan, double* ms_by_tag, int64_t* count_by_tag);

/* Dictionary: d is (n_atoms, atom_size) row-major.  The plan keeps
 * x / (||x||_2 + 1e-8) per row -- modules/normalization.py:4-6, applied by the
 * reference at every entry (modules/matchingpursuit.py:254) -- and derives the
 * atom-pair spectra (and the Gram table in GRAM mode).  The caller's buffer is
 * not modified and may be freed once `stream` has passed this call.  A
 * 128-bit fingerprint of `d` (two independently keyed sums of a full-avalanche mix of every (index, value))
 * is compared ON THE DEVICE with the previous call's: when
 * the bits are unchanged the table-building kernels return at once, so callers
 * that keep the reference's habit of passing the dictionary with every call
 * do not pay for it, and nothing synchronises with the host. */
int mpb200_plan_set_dictionary(mpb200_plan_t plan, const float* d, void* stream);
/* Same without the normalisation: the atoms are used as given -- what the
 * correlation helpers modules/conv.py:4-9 (torch_conv) and :11-53 (fft_convolve)
 * do, and mp.py's un-normalised atoms (mp.py:43-48). */
int mpb200_plan_set_dictionary_raw(mpb200_plan_t plan, const float* d, void* stream);
/* Copy of the normalised dictionary (n_atoms, atom_size). */
int mpb200_plan_get_unit_dictionary(mpb200_plan_t plan, float* out, void* stream);

/* Greedy pursuit -- replaces the loop of modules/matchingpursuit.py:269-328
 * (sparse_code) and :87-120 (sparse_feature_map).
 *   signal        (batch, n_samples), not modified
 *   residual_out  (batch, n_samples), receives signal - sum of selected atoms
 *   atom_out/pos_out/val_out  (batch, n_steps) row-major: step s of signal b at [b*n_steps + s]
 * Exactly n_steps events per signal (no early stop); value is the signed
 * correlation at the winner; ties go to the lowest (atom, position). */
int mpb200_sparse_code(mpb200_plan_t plan, const float* signal, int batch, int n_steps,
                       float* residual_out, int32_t* atom_out, int32_t* pos_out, float* val_out,
                       void* stream);
/* Same with HOST buffers (pinned or pageable): copies in, runs, copies out and
 * synchronises `stream` before returning.  This is the end-to-end entry the
 * Python drop-ins use for CPU tensors. */
int mpb200_sparse_code_host(mpb200_plan_t plan, const float* signal_host, int batch, int n_steps,
                            float* residual_out_host, int32_t* atom_out_host, int32_t* pos_out_host,
                            float* val_out_host, void* stream);

/* Dense correlation map fm[b,k,t] = sum_i pad(signal)[b,t+i] * d_unit[k,i],
 * (batch, n_owned_atoms, n_samples) -- replaces F.pad+F.conv1d+crop at
 * modules/matchingpursuit.py:275-277, modules/conv.py:4-9 (torch_conv) and
 * the full-product branch of modules/conv.py:11-53 (fft_convolve); it is also
 * the `compute_feature_map` callback seam (modules/matchingpursuit.py:272-273). */
int mpb200_correlate(mpb200_plan_t plan, const float* signal, int batch, float* fm_out, void* stream);

/* Step-wise interface (atom sharding across GPUs, and callers that need to
 * look at every step):
 *   begin       load signals, run the first full pass
 *   local_best  this plan's best (value, global atom, position) per signal -> best[batch]
 *   apply       subtract winner[b] (an atom this plan may not own: the whole
 *               dictionary is replicated) from the residual and refresh this
 *               plan's map/block maxima in the +-A window
 *   residual    copy out the current residual
 * Between local_best and apply the caller reduces the records across ranks
 * (max value, then lowest atom, then lowest position). */
int mpb200_begin(mpb200_plan_t plan, const float* signal, int batch, void* stream);
int mpb200_local_best(mpb200_plan_t plan, mpb200_best* best, void* stream);
int mpb200_apply(mpb200_plan_t plan, const mpb200_best* winner, void* stream);
int mpb200_residual(mpb200_plan_t plan, float* residual_out, void* stream);
/* Reduce `n_ranks` candidate lists (rank-major: cand[r*batch + b]) to the
 * global winner per signal with the reference tie-break; used after an
 * all-gather of mpb200_local_best records. */
int mpb200_reduce_best(const mpb200_best* cand, int n_ranks, int batch, mpb200_best* winner, void* stream);

/* Atom sharding with the exchange fused into the pursuit (no collective library in the loop).  Every rank
 * creates a mailbox (2 x max_batch x world slots of 32 bytes, plain device memory) and the ranks trade the
 * 64-byte CUDA IPC handles out of band (any transport: the Python layer uses one torch.distributed
 * all-gather at set-up).  After mpb200_exchange_connect, mpb200_sparse_code works on the atom-sharded plan:
 * in every iteration the kernel that applies the winner first writes this rank's candidate (value, atom,
 * position, each in an 8-byte word with the iteration's sequence number) straight into every peer's mailbox
 * over NVLink, polls its own mailbox for the `world` records of this iteration, and reduces them with the
 * reference tie-break (max value, then lowest atom, then lowest position).  All ranks must issue the same
 * calls in the same order.  A record that does not arrive within 20 s sets a flag (mpb200_exchange_status)
 * instead of hanging.  No reference counterpart (the reference is single-device).
 *   connect        handles = world x 64 bytes in rank order (this rank's own entry is ignored)
 *   connect_local  same process, several plans (tests): mailboxes[r] = pointer from mpb200_exchange_mailbox */
int mpb200_exchange_create(mpb200_plan_t plan, int world, int rank, unsigned char* handle_out /* 64 bytes or NULL */);
int mpb200_exchange_connect(mpb200_plan_t plan, const unsigned char* handles);
int mpb200_exchange_mailbox(mpb200_plan_t plan, void** mailbox);
int mpb200_exchange_connect_local(mpb200_plan_t plan, void* const* mailboxes);
int mpb200_exchange_status(mpb200_plan_t plan, int* timed_out);
/* Collective teardown, first half: waits for the plan's device work and unmaps the peers' mailboxes.  All ranks call
 * it, synchronise among themselves (any transport), and only then destroy their plans -- freeing a mailbox that a
 * peer still has mapped through CUDA IPC is undefined.  The plan is a plain un-connected sharded plan afterwards. */
int mpb200_exchange_disconnect(mpb200_plan_t plan);

/* Selection on a DENSE map fm (batch, n_atoms, n_samples) that the caller
 * already holds (a `compute_feature_map` callback result, or
 * mpb200_correlate output handed to a per-step visitor): signed maximum per
 * signal, first flat index on ties -- torch.max over fm.reshape(batch, -1) at
 * modules/matchingpursuit.py:298-303; also the top-1 of sparsify2,
 * modules/sparse.py:64.  best[b].atom = atom_offset + row of the maximum. */
int mpb200_select_dense(const float* fm, int batch, int n_atoms, int n_samples, int atom_offset,
                        mpb200_best* best, void* stream);
/* Same, with the reference's opt-in local contrast normalisation: the argmax
 * runs on fm - avg_pool2d(fm, (9,9), stride 1, zero padding 4) over the
 * (atom, time) plane and the reported value is the RAW map value at that
 * index -- modules/matchingpursuit.py:286-296. */
int mpb200_select_lcn(const float* fm, int batch, int n_atoms, int n_samples, int atom_offset,
                      mpb200_best* best, void* stream);
/* residual[b, p : p+atom_size] -= winner[b].value * d_unit[winner[b].atom],
 * p = winner[b].position, truncated at the right edge -- the residual update
 * of modules/matchingpursuit.py:326-328 (and :108-120). */
int mpb200_subtract(float* residual, int batch, int n_samples, const float* d_unit, int n_atoms, int atom_size,
                    const mpb200_best* winner, void* stream);

/* Decode: out[b, pos : pos+atom_size] += val * d_unit[atom], truncated at the
 * right edge -- replaces scatter_segments, modules/matchingpursuit.py:20-58
 * (single-channel branch :48).  out is (batch, n_samples) and is accumulated
 * into (zero it first for a fresh decode).  Events: n_events entries of
 * (atom, batch index, position, value), summed in list order per sample.
 * row_offsets: NULL, or batch+1 offsets when the events are sorted by batch
 * index (events of signal b are [row_offsets[b], row_offsets[b+1])). */
int mpb200_scatter_add(float* out, int batch, int n_samples, const float* d_unit, int n_atoms, int atom_size,
                       const int32_t* atom, const int32_t* batch_index, const int32_t* pos, const float* val,
                       const int32_t* row_offsets, int n_events, void* stream);
/* Same with caller-supplied scaled atoms: out[row_index[e], pos[e] : +atom_size] += rows[e, :]
 * (rows is (n_events, atom_size)) -- scatter_segments fed with arbitrary event
 * tuples (modules/matchingpursuit.py:41-52, used by dictionary_learning_step
 * :408-415 and BandSpec.decode, modules/multibanddict.py:265-266).  out is
 * (n_rows, n_samples); for the reference's one-channel-per-event mode (:50) the
 * caller passes row = batch * channels + channel. */
int mpb200_scatter_rows(float* out, int n_rows, int n_samples, const float* rows, int atom_size,
                        const int32_t* row_index, const int32_t* pos, const int32_t* row_offsets, int n_events,
                        void* stream);
/* scaled[e, :] = val[e] * d_unit[atom[e], :]  -- the `a` member of the
 * reference's event tuples (modules/matchingpursuit.py:305, 315). */
int mpb200_gather_atoms(float* scaled, const float* d_unit, int n_atoms, int atom_size,
                        const int32_t* atom, const float* val, int n_events, void* stream);

/* Dictionary-learning atom update -- the loop of modules/matchingpursuit.py:391-417 as one launch.  The events of a
 * coding pass (mpb200_sparse_code + mpb200_gather_atoms) are given GROUPED BY ATOM in the reference's first-seen
 * order: group g holds events [group_offsets[g], group_offsets[g+1]) of atom group_atom[g], inside a group in
 * step-major, batch-minor order (:261, :321).  For every group, in order: the group's scaled atoms ev_rows[e]
 * (atom_size samples each) are added back to `running` at (ev_batch[e], ev_pos[e]), the new atom is the unit-normed
 * (modules/normalization.py:4-6) sum of the running-signal segments under the group's events (zero beyond the
 * signal), it replaces row group_atom[g] of d_unit, and new_atom * ||ev_rows[e]|| is subtracted at every event.
 *   running  (batch, n_samples) in/out: starts as a copy of the SIGNAL (:367 -- not the coding residual)
 *   d_unit   (n_atoms, atom_size) in/out: the unit-normed dictionary the events were coded with
 * The caller applies the final unit_norm of :417 (mpb200_unit_norm). */
int mpb200_dictionary_update(float* running, int batch, int n_samples, float* d_unit, int n_atoms, int atom_size,
                             const int32_t* group_offsets, const int32_t* group_atom, int n_groups,
                             const int32_t* ev_batch, const int32_t* ev_pos, const float* ev_rows, int n_events,
                             void* stream);

/* Atoms longer than a plan can take (mpb200_plan_create: atom_size <= MPB200_MAX_PLAN_ATOM; the reference runs
 * 4096, 8192 and 16384 samples in experiments/archive/e_2023_3_8/experiment.py:352-358 and
 * e_2023_12_18/experiment.py:22-24) are correlated as n_parts consecutive parts of part_len samples: the caller
 * makes a plan for the (n_atoms * n_parts, part_len) dictionary of parts (part p of atom k in row k*n_parts + p, the
 * last part zero padded), runs mpb200_correlate, and this call folds the parts' maps into the long atoms' map:
 *     fm_out[b, k, t] = sum_p sub_map[b, k*n_parts + p, t + p*part_len]      (terms beyond n_samples are zero)
 * which equals F.conv1d of modules/matchingpursuit.py:275-277 for the long atoms.  sub_map is
 * (batch, n_atoms * n_parts, n_samples), fm_out (batch, n_atoms, n_samples). */
#define MPB200_MAX_PLAN_ATOM 2560
int mpb200_fold_parts(const float* sub_map, int batch, int n_atoms, int n_parts, int part_len, int n_samples,
                      float* fm_out, void* stream);

/* The same dense map as mpb200_correlate, computed on the TENSOR CORES instead of by FFT (BASELINE.json north_star
 * (1)): a Toeplitz/Hankel GEMM with tcgen05.mma (kind::tf32, accumulators in tensor memory) and split-precision
 * 3xTF32 so that fp32 argmax parity holds; the Hankel operand is expanded in shared memory.  The atoms are used as
 * given (d is (n_atoms, atom_size), not normalised); fm_out is (batch, n_atoms, n_samples).  The plans never choose
 * this route on their own -- it is exported so that the two routes can be measured against each other with the
 * tensor-pipe and HBM counters (profiles/r2_gemm_vs_fft.md).  spin_blocks > 0 turns the call into a tensor-pipe
 * peak probe: nothing is written, and every CTA issues that many extra 128x256x32 3xTF32 MMA blocks. */
int mpb200_correlate_gemm(const float* signal, int batch, int n_samples, const float* d, int n_atoms, int atom_size,
                          float* fm_out, int spin_blocks, void* stream);

/* y = x / (||x||_2 + eps) per row -- modules/normalization.py:4-6. */
int mpb200_unit_norm(const float* x, float* y, int rows, int cols, float eps, void* stream);

/* N-ary zero-padded FFT convolution -- replaces modules/fft.py:23-35 (and the
 * identical arithmetic of modules/transfer.py:548-569): every operand row of n
 * samples is padded to 2n, the spectra are multiplied, the product is inverted
 * at length 2n and cropped to n.
 *   operands      HOST array of n_ops (<= 4) DEVICE pointers, operand i is (operand_rows[i], n)
 *   row_maps      HOST array of n_ops DEVICE pointers (or NULL entries = identity): row_maps[i][r]
 *                 is the row of operand i that output row r uses (how the caller expresses
 *                 broadcasting over leading dimensions); rows_out entries each
 *   conj_mask     bit i set: operand i's spectrum is conjugated (correlation instead of convolution)
 *   scale         applied to the result (1 for norm=None; the caller folds norm="ortho"/"forward" in)
 *   out           (rows_out, n)
 * Internally a power-of-two transform of at least the full linear length is used and the result is
 * wrapped onto period 2n, so 3- and 4-operand products alias exactly as the reference's do. */
int mpb200_fft_convolve(const float* const* operands, const int32_t* const* row_maps,
                        const int32_t* operand_rows, int n_ops, int rows_out, int n, int conj_mask,
                        float scale, float* out, void* stream);

/* y = irfft(mask(rfft(pad(x, L)))) per row, where the mask keeps the n_bins bins bin0, bin0+bin_step, ...
 * (<= L/2) and zeroes the rest -- the spectral mask of modules/conv.py:24-29, fft_convolve(approx=slice),
 * whose transform length is L = n_samples + atom_size (even; not a power of two in general, so this is a
 * direct DFT over the kept bins).  The reference's band-limited correlation map is then the ordinary
 * correlation (mpb200_correlate on a plan of L samples, first n_samples columns) of y with the atoms.
 *   x (rows, n), y (rows, L). */
int mpb200_band_limit(const float* x, int rows, int n, int L, int bin0, int bin_step, int n_bins, float* y,
                      void* stream);

/* out = irfft(keep bins [bin_lo, bin_hi) of rfft(x, norm="ortho"), n=n_out, norm="ortho") per row;
 * x is (rows, n_in), out (rows, n_out), both lengths powers of two >= 256 -- the band split of
 * modules/decompose.py:5-33 (fft_frequency_decompose) and the zero-stuffing resample of :36-73
 * (fft_resample; its tukey(alpha=0) window is identically one). */
int mpb200_spectral_band(const float* x, int rows, int n_in, float* out, int n_out, int bin_lo, int bin_hi,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPB200_H */
