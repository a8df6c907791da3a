/*
 * amofb.h -- C ABI of libamofb.so, the B200 (sm_100a) implementation of aMOF's
 * frame-parallel structural-analysis hot path.
 *
 * The reference (coudertlab/amof) is pure Python and has NO plugin / FFI interface of its own
 * (SURVEY.md 8(b)); its hot arithmetic is delegated to asap3's C++ extension and to ase/numpy.
 * This header is therefore the NEW boundary: each entry point names the reference interface it
 * replaces.  The Python classes in amof_b200/ (same names and signatures as amof.rdf / amof.cn /
 * amof.bad / amof.msd) are its only in-tree caller, through ctypes; INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative AMOFB_ERR_* code;
 *     amofb_last_error(ctx) gives the message.  No C++ exception crosses the ABI.
 *   - positions are double[n_frames][n_atoms][3] (Angstrom), cells double[n_frames][3][3] with the
 *     lattice vectors as rows -- exactly ase.Atoms.get_positions() / get_cell() stacked per frame.
 *     All three directions are periodic (the reference assumes pbc=True throughout).
 *   - species are small indices 0..n_species-1 (uint8), one per atom, constant over the trajectory.
 *   - the caller owns every pointer it passes; the library owns device memory, pinned staging and
 *     streams.  "push" calls enqueue work and may return before it has run; "finish"/"sync" wait.
 *   - one ctx per GPU; a ctx is not thread-safe, distinct ctxs are independent.
 *   - there is NO CPU fallback: without a CUDA device amofb_create fails.
 *   - integer outputs (histograms, neighbour counts) are bit-exact with oracle/amof_oracle.c;
 *     normalisation to g(r), densities and means stays in the caller (fp64 numpy).
 */
#ifndef AMOFB_H
#define AMOFB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMOFB_OK 0
#define AMOFB_ERR_ARG (-1)      /* bad argument / call order            */
#define AMOFB_ERR_CUDA (-2)     /* CUDA runtime error                   */
#define AMOFB_ERR_STATE (-3)    /* begin/push/finish called out of turn */
#define AMOFB_ERR_GEOMETRY (-4) /* singular cell, cutoff precondition, neighbour overflow */
#define AMOFB_ERR_MEMORY (-5)

#define AMOFB_MAX_SPECIES 16
#define AMOFB_BAD_MAX_CN 32     /* a centre with more B-neighbours than this is a geometry error */

typedef struct amofb_ctx amofb_ctx;

/* ---- lifetime ------------------------------------------------------------------------------ */
int amofb_create(int device, amofb_ctx **out);
int amofb_destroy(amofb_ctx *ctx);
const char *amofb_last_error(const amofb_ctx *ctx); /* valid until the next call on ctx */
const char *amofb_version(void);
/* Block until everything enqueued on ctx has finished (used by benchmarks to close a timed region). */
int amofb_sync(amofb_ctx *ctx);
/* Block until every host->device copy enqueued so far has finished (kernels may still be running): after this the
 * caller may overwrite the pinned buffers it passed to earlier push/load calls. */
int amofb_sync_copies(amofb_ctx *ctx);
/* Number of kernels this ctx has launched since creation (bench.py reports it as gpu_launches). */
int64_t amofb_launch_count(const amofb_ctx *ctx);
/* Debug aid.  With AMOFB_GUARD=1 in the environment at amofb_create every pooled device block carries 4 KiB of canary bytes on
 * either side, compared when the block is returned; this is how many blocks were found written out of bounds so far (each is
 * also reported on stderr).  -1 when the guard is off. */
int64_t amofb_guard_violations(const amofb_ctx *ctx);
/* Conventions of the upstream packages that cannot be read off their sources here (asap3 / ase are not on disk,
 * SURVEY.md 8(c) U1-U6) are options with the oracle's pin as default:
 *   AMOFB_OPT_RDF_BIN_RULE  0: bin = (int)(d / (rMax/nBins))   (default, pin U1)
 *                           1: bin = (int)(d * (nBins/rMax))
 * Set between analyses (AMOFB_ERR_STATE while a pair analysis is open). */
#define AMOFB_OPT_RDF_BIN_RULE 1
int amofb_set_option(amofb_ctx *ctx, int option, int value);
/* Sum of CUDA-event durations (ms) and launch count of the pair kernel since the last call with
 * reset != 0; timing is only collected after amofb_set_profiling(ctx, 1). */
int amofb_set_profiling(amofb_ctx *ctx, int enabled);
int amofb_pair_kernel_time(amofb_ctx *ctx, double *total_ms, int64_t *launches, int reset);

/* Device-side stopwatch for benchmarks: amofb_timer_mark records CUDA event `slot` (0..7) on the compute stream,
 * amofb_timer_elapsed waits for both events and returns the milliseconds between them. */
int amofb_timer_mark(amofb_ctx *ctx, int slot);
int amofb_timer_elapsed(amofb_ctx *ctx, int from_slot, int to_slot, double *ms);

/* Page-locked host buffers for callers that want zero-copy streaming (bench.py, the Python classes). */
int amofb_host_alloc(amofb_ctx *ctx, uint64_t bytes, void **out);
int amofb_host_free(amofb_ctx *ctx, void *ptr);
/* Device buffers for device-resident trajectories (amofb_*_push_device). */
int amofb_device_alloc(amofb_ctx *ctx, uint64_t bytes, void **out);
int amofb_device_free(amofb_ctx *ctx, void *ptr);
int amofb_memcpy_h2d(amofb_ctx *ctx, void *dst_device, const void *src_host, uint64_t bytes);
int amofb_memcpy_d2h(amofb_ctx *ctx, void *dst_host, const void *src_device, uint64_t bytes);

/* ---- pair analysis: partial RDF histograms and cutoff neighbour counts ---------------------
 * Replaces, per frame,
 *   asap3.analysis.rdf.RadialDistributionFunction(atoms, rMax, nBins) / .atoms = frame / .update()
 *       (call sites /root/reference/amof/rdf.py:87-93 and :181)
 *   ase.neighborlist.neighbor_list('ij', atoms, cutoff_dict) + the counting loops
 *       (/root/reference/amof/atom.py:72-87, amof/cn.py:58-74)
 * in ONE pass over the trajectory.
 *
 * begin : nbins > 0 enables the RDF histogram with bin width rmax/nbins (bin = (int)(d / (rmax/nbins)),
 *         counted iff < nbins); cn_cutoff != NULL (double[n_species][n_species], symmetric, 0 = pair not
 *         listed) enables per-frame neighbour counts with the strict test d < cutoff[Zi][Zj].
 * push  : n_frames frames in host memory (pinned memory is copied asynchronously, pageable memory is
 *         staged).  push_device: positions already in device memory (cells still on the host); the data must be
 *         complete when the call is made (the library reads it on its own stream) and stay valid until finish/sync.
 * finish: hist       uint64[n_species][n_species][nbins]  directed pair counts, summed over frames (or NULL)
 *         cn_counts  uint64[cn_frames][n_species][n_species] directed neighbour pairs per frame (or NULL);
 *                    cn_frames must equal the number of frames pushed
 *         n_frames_out, volume_sum_out: frames seen and the sum of their cell volumes
 *         finish ends the accumulation; begin may be called again on the same ctx.
 * take  : finish's outputs for the frames pushed since begin (or since the last take), after which the accumulators are
 *         empty and the analysis stays OPEN: one histogram per frame, as amof.rdf.CoordinationNumber builds a fresh
 *         RadialDistributionFunction for every frame (/root/reference/amof/rdf.py:181-186), without re-planning.
 */
int amofb_pair_begin(amofb_ctx *ctx, int n_atoms, int n_species, const uint8_t *species,
                     double rmax, int nbins, const double *cn_cutoff);
int amofb_pair_push(amofb_ctx *ctx, int n_frames, const double *pos, const double *cell);
int amofb_pair_push_device(amofb_ctx *ctx, int n_frames, const double *pos_device, const double *cell);
int amofb_pair_take(amofb_ctx *ctx, uint64_t *hist, uint64_t *cn_counts, int64_t cn_frames,
                    int64_t *n_frames_out, double *volume_sum_out);
int amofb_pair_finish(amofb_ctx *ctx, uint64_t *hist, uint64_t *cn_counts, int64_t cn_frames,
                      int64_t *n_frames_out, double *volume_sum_out);

/* The accumulator trio SURVEY.md 8(b) proposed, as thin aliases of the pair analysis:
 * amofb_rdf_* = histogram only, amofb_cn_* = neighbour counts only. */
int amofb_rdf_begin(amofb_ctx *ctx, int n_atoms, int n_species, const uint8_t *species, double rmax, int nbins);
int amofb_rdf_push(amofb_ctx *ctx, int n_frames, const double *pos, const double *cell);
int amofb_rdf_finish(amofb_ctx *ctx, uint64_t *hist, int64_t *n_frames_out, double *volume_sum_out);
int amofb_cn_begin(amofb_ctx *ctx, int n_atoms, int n_species, const uint8_t *species, const double *cn_cutoff);
int amofb_cn_push(amofb_ctx *ctx, int n_frames, const double *pos, const double *cell);
int amofb_cn_finish(amofb_ctx *ctx, uint64_t *cn_counts, int64_t cn_frames);

/* ---- bond-angle distributions ---------------------------------------------------------------
 * Replaces amof.atom.get_neighborlist + Bad.bad_BAB / BadByCn.bad_BAB + np.histogram
 * (/root/reference/amof/bad.py:70-114,160 and :192-239,293): for every centre a of species A, all unordered
 * pairs of its B-neighbours (neighbour list under cutoff[n_species][n_species]) give one angle in degrees,
 * histogrammed on the edges k*dtheta, k = 0..nbins, split by the centre's number of B-neighbours.
 *
 * triples: int[n_triples][2] = (A, B) species indices, -1 = "X" (any species).
 * finish : hist uint64[n_triples][AMOFB_BAD_MAX_CN+1][nbins]; dropped uint64[n_triples] counts angles that
 *          np.histogram would drop (NaN from |cos| > 1 by rounding).
 * Precondition (AMOFB_ERR_GEOMETRY otherwise): max cutoff < half the smallest perpendicular cell height.
 */
int amofb_bad_begin(amofb_ctx *ctx, int n_atoms, int n_species, const uint8_t *species,
                    const double *cutoff, int n_triples, const int *triples, double dtheta, int nbins);
int amofb_bad_push(amofb_ctx *ctx, int n_frames, const double *pos, const double *cell);
int amofb_bad_push_device(amofb_ctx *ctx, int n_frames, const double *pos_device, const double *cell);
int amofb_bad_finish(amofb_ctx *ctx, uint64_t *hist, uint64_t *dropped, int64_t *n_frames_out);

/* ---- explicit neighbour list of one frame -------------------------------------------------------
 * Replaces amof.atom.get_neighborlist (/root/reference/amof/atom.py:72-87):
 *   nl_i, nl_j = ase.neighborlist.neighbor_list('ij', atom, cutoff_dict), regrouped as one list of j per atom i.
 * The analyses above never build the list (the search is fused with the counting and with the angle enumeration);
 * this pair of calls serves the callers that need the list itself (amof.ring, amof.coordination: SURVEY.md 8(f)).
 * A pair (i, j, image) is listed iff d < cutoff[Zi][Zj] (strict, 0 = pair not listed); the zero-shift self pair is
 * skipped; j appears once per periodic image under the cutoff, as ase lists it.
 *
 * count : one frame (pos double[n_atoms][3], cell double[9]); offsets int64[n_atoms + 1] receives the row starts
 *         (offsets[n_atoms] = number of directed pairs).  The search state stays open for fill.
 * fill  : neighbors int32[capacity], capacity >= offsets[n_atoms]: row i = neighbors[offsets[i] .. offsets[i+1]),
 *         ORIGINAL atom indices in ascending order (ase's own order inside a row is unspecified).  Ends the search.
 * fill_ex: the same plus, per listed pair and in the same order, ase's quantities 'd' and 'S' -- what
 *         pymatgen's Structure.get_neighbor_list hands amof.coordination (/root/reference/amof/coordination/core.py:62,181):
 *         distances double[capacity] (or NULL) = |D|, shifts int32[capacity][3] (or NULL) = the integer image S with
 *         D = p_j - p_i + S.cell for the positions as given (not wrapped).  Rows are ordered by (j, S).
 */
int amofb_neigh_count(amofb_ctx *ctx, int n_atoms, int n_species, const uint8_t *species, const double *cutoff,
                      const double *pos, const double *cell, int64_t *offsets);
int amofb_neigh_fill(amofb_ctx *ctx, int32_t *neighbors, int64_t capacity);
int amofb_neigh_fill_ex(amofb_ctx *ctx, int32_t *neighbors, double *distances, int32_t *shifts, int64_t capacity);

/* ---- mean-squared displacement --------------------------------------------------------------
 * Replaces WindowMsd.compute_msd / compute_msd_of_m and trajectory.get_delta_pos
 * (/root/reference/amof/msd.py:186-268, amof/trajectory.py:285-303).
 *
 * The trajectory is resident on the device for the whole computation (unwrapping is a scan over frames and
 * window pairs span up to half the trajectory), so MSD is ATOM-sharded across GPUs: every rank holds all
 * frames of its own atoms, and the only exchanges are the per-frame centre-of-mass sums and the final sums.
 *
 * begin      : n_frames x n_atoms (local atoms), masses[n_atoms], species[n_atoms], cell[n_frames][9].
 * load       : copy frames [first, first+count) of the local atoms, host double[count][n_atoms][3].
 * load_device: same from device memory.
 * unwrap     : optional (WindowMsd(unwrap=True), msd.py:222-230): positions <- cumulative wrapped displacements.
 * com_sums   : double[n_frames][4] = (sum m*x, sum m*y, sum m*z, sum m) over the LOCAL atoms.
 * set_com    : double[n_frames][3] global centre of mass per frame (after the caller's allreduce).
 * window     : sums[n_species][n_window] = sum over local atoms of species s and over k = m+1..T-1 of
 *              |R_k - R_{k-m}|^2, R = COM-removed positions rebuilt from wrapped displacements.
 *              The caller divides by N_species*(T-m) (quirk Q4 of SURVEY.md) after its allreduce.
 * direct     : DirectMsd.compute_species_msd (msd.py:81-105), orthogonal cells only:
 *              sums[n_species][n_frames] = sum over local atoms of |r_t - r_0|^2.
 * get_positions: copy the positions (unwrapped if unwrap was called) back, host double[n_frames][n_atoms][3].
 *              Call it before window: window may rewrite the trajectory in place (running sums of displacements).
 */
int amofb_msd_begin(amofb_ctx *ctx, int n_frames, int n_atoms, const double *masses, const uint8_t *species,
                    int n_species, const double *cell);
int amofb_msd_load(amofb_ctx *ctx, int first_frame, int count, const double *pos);
int amofb_msd_load_device(amofb_ctx *ctx, int first_frame, int count, const double *pos_device);
int amofb_msd_unwrap(amofb_ctx *ctx);
int amofb_msd_com_sums(amofb_ctx *ctx, double *sums);
int amofb_msd_set_com(amofb_ctx *ctx, const double *com);
 /* Streaming path for WindowMsd(unwrap=False): everything that is a pass over the positions happens on the way in.
 * Per slab of consecutive frames, in frame order from frame 0 to n_frames:
 *   slab_sums  : the slab (host double[count][n_atoms][3], count <= amofb_msd_slab_frames; the _device variant reads a
 *                device pointer in place, any count, valid until the commit) -> sums double[count][4] =
 *                (sum m*x, sum m*y, sum m*z, sum m) over the LOCAL atoms of each frame;
 *   slab_commit: com double[count][3], the GLOBAL centre of mass of those frames (after the caller's all-reduce):
 *                centre-of-mass shift (msd.py:235-237), displacement wrap with the cell of the earlier frame and running
 *                sum (trajectory.py:285-303), written into the atom-major store.  Returns once enqueued.
 * After the last commit amofb_msd_window takes window lengths 0, D, 2D, ... (what WindowMsd always asks for,
 * msd.py:176-178); load / unwrap / com_sums / set_com / direct / get_positions belong to the other path. */
int amofb_msd_slab_frames(amofb_ctx *ctx, int *frames);
int amofb_msd_slab_sums(amofb_ctx *ctx, int first_frame, int count, const double *pos, double *sums);
int amofb_msd_slab_sums_device(amofb_ctx *ctx, int first_frame, int count, const double *pos_device, double *sums);
/* The same in two halves, so that the next slab's sums are already enqueued while the caller turns the previous
 * slab's sums into a centre of mass: begin enqueues (at most two slabs may await their commit), wait returns the sums
 * of the oldest slab whose sums were not fetched yet.  slab_sums == begin + wait. */
int amofb_msd_slab_sums_begin(amofb_ctx *ctx, int first_frame, int count, const double *pos);
/* strided: frame k of the slab starts at pos + k * frame_stride doubles (frame_stride >= 3 * n_atoms): the local atoms are a
 * column block of wider frames (an atom-sharded rank reading a whole-frame trajectory), copied without host-side packing */
int amofb_msd_slab_sums_begin_strided(amofb_ctx *ctx, int first_frame, int count, const double *pos, int64_t frame_stride);
int amofb_msd_slab_sums_begin_device(amofb_ctx *ctx, int first_frame, int count, const double *pos_device);
int amofb_msd_slab_sums_wait(amofb_ctx *ctx, double *sums);
int amofb_msd_slab_commit(amofb_ctx *ctx, const double *com);
int amofb_msd_window(amofb_ctx *ctx, int n_window, const int *window, double *sums);
int amofb_msd_direct(amofb_ctx *ctx, double *sums);
int amofb_msd_get_positions(amofb_ctx *ctx, double *pos);
int amofb_msd_end(amofb_ctx *ctx);

/* ------------------------------------------------------------------------------------------------------------------
 * Trajectory text ingest (host code, no device work): the step in front of every analysis, replacing the per-frame
 * Python parsing of ase.io.read(filename, index, 'xyz') (/root/reference/amof/trajectory.py:48-60,193-228).
 *
 * xyz_parse: n_frames XYZ / extended-XYZ frames out of a text buffer; frame k occupies text[frame_off[k] .. frame_off[k+1])
 *            and is a count line, a comment line and n_atoms atom lines "Symbol ... x y z ...", x being whitespace-separated
 *            column pos_col (1 in plain XYZ).  positions double[n_frames][n_atoms][3] receives the coordinates, every decimal
 *            string converted with correct rounding (the value Python's float() gives).  symbols is char[n_atoms][8],
 *            NUL-padded: filled from the first frame when symbols_known == 0, otherwise every frame is checked against it.
 *            Frames are spread over `threads` host threads (<= 0: one per core, at most 16).
 *            AMOFB_ERR_ARG: malformed or truncated frame, or an atom order that changes; *bad_frame (may be NULL) says which.
 * xyz_index: where frames start.  Scans a block of the file (text[0 .. len), whose first byte is byte `base` of the file and
 *            which `lines_before` complete lines precede) and writes the file offset of every line whose number is a multiple of
 *            `period` (= n_atoms + 2) into starts[capacity]; *n_starts = how many there are (AMOFB_ERR_MEMORY if more than
 *            capacity: nothing beyond capacity is written), *n_lines = newlines seen in the block.
 */
int amofb_xyz_index(const char *text, int64_t len, int64_t lines_before, int64_t period, int64_t base, int64_t *starts,
                    int64_t capacity, int64_t *n_starts, int64_t *n_lines);
int amofb_xyz_parse(const char *text, const int64_t *frame_off, int n_frames, int n_atoms, int pos_col, char *symbols,
                    int symbols_known, double *positions, int threads, int *bad_frame);

#ifdef __cplusplus
}
#endif
#endif /* AMOFB_H */
