// NOT PART OF THE LIBRARY.  Register version of the MSD ingest kernel (k_msd_slab_commit), measured in round 2 and not kept:
// 9.1 ms (as first written) and 9.6-9.8 ms (rows made branch-free for ILP; the 16-row array went to local memory) against
// 8.85 ms of the block kernel, per 100 000 atoms x 5 000 frames.  ncu: 75 instructions per (column, frame) -- 12 SHFL for the
// 3x3 products, 12 IMAD -- 50 % issue utilisation, 24 warps per SM at 80 registers.  Bit-identical results (it was tested
// against the block kernel for the three cell modes before it left the tree).  To build it again include this file after
// msd.cuh's wrap helpers and launch it as msd_host.inl launches k_msd_slab_commit.
// Register version (round 2; ncu of the kernel above: 7 barriers per round, the serial scan -- half of the block idle -- holds
// 37 % of the stall samples, 3.2 TB/s).  A lane owns one (atom, component) COLUMN for the whole slab: previous position and
// running sum stay in registers.  A warp holds 10 atoms (30 lanes: the three components of an atom are neighbouring lanes, so
// the wrap's 3x3 products take the other components by shuffle; 2 lanes idle), a block 12 warps = 120 atoms.  Per round of
// REGF frames: REGF independent loads per lane (240-byte runs per warp and frame, the warps of a block back to back), then
// shift -> difference -> wrap -> running sum in registers, the sums into a padded tile, ONE barrier, and the transposed
// write-out (half a warp per column: 128-byte runs of the atom-major store).  Two tiles alternate, so the write-out of a round
// overlaps the loads of the next.  Expressions and order are those of wrap_disp / wrap_disp_diag: the same bits.
#define REGF 16
#ifndef REG_ILP
#define REG_ILP 4                         // rows whose wrap chains are in flight together
#endif
#define REG_THREADS 384
#define REG_ATOMS (10 * (REG_THREADS / 32))
#define REG_COLS (3 * REG_ATOMS)
#define REG_LD (REG_COLS + 1)
#define REG_COM_MAX 256                   // frames per host slab (stage_frames is at most 256); longer device slabs take the kernel above
#define REG_SMEM (sizeof(double) * (2 * REGF * REG_LD + 3 * (REG_COM_MAX + REGF)))      // the centre-of-mass rows are read REGF at a time, past the slab's end too
template <int CELL>      // 0 = one cell per frame, 1 = the same cell in every frame, 2 = the same orthorhombic cell
__global__ void __launch_bounds__(REG_THREADS, 2) k_msd_slab_commit_reg(const double *__restrict__ slab, double *__restrict__ P,
                                                                        const MsdGeom *__restrict__ geom, const double *__restrict__ com,
                                                                        double *__restrict__ carry, int n, int Tp, int first, int count) {
    extern __shared__ __align__(16) double reg_sm[];
    double *s_com = reg_sm + 2 * REGF * REG_LD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a0 = blockIdx.x * REG_ATOMS, na = min(REG_ATOMS, n - a0), ncol = 3 * na;
    const int comp = lane % 3, base = lane - comp;            // lanes 30, 31: comp 0 / 1 of a phantom atom, never `mine`
    const int col = 30 * warp + lane;
    const bool mine = lane < 30 && col < ncol;
    const size_t cat = (size_t)(a0 + col / 3) * 6 + (size_t)comp;
    double prev = (mine && first > 0) ? carry[cat] : 0.0, run = (mine && first > 0) ? carry[cat + 3] : 0.0;
    const double shift = (0.0 - 0.5) - 1e-7;
    double i0 = 0.0, i1 = 0.0, i2 = 0.0, c0 = 0.0, c1 = 0.0, c2 = 0.0;     // inverse column `comp`, cell column `comp`
    if (CELL != 0) {
        i0 = geom[0].inv[comp]; i1 = geom[0].inv[3 + comp]; i2 = geom[0].inv[6 + comp];
        c0 = geom[0].cell[comp]; c1 = geom[0].cell[3 + comp]; c2 = geom[0].cell[6 + comp];
    }
    for (int i = tid; i < 3 * (REG_COM_MAX + REGF); i += REG_THREADS) s_com[i] = i < 3 * count ? com[i] : 0.0;
    __syncthreads();
    const double *src = slab + (size_t)a0 * 3 + col;
    const size_t fstride = (size_t)n * 3;
    int buf = 0;
    for (int k0 = 0; k0 < count; k0 += REGF, buf ^= 1) {
        const int nr = min(REGF, count - k0);
        double *tile = reg_sm + (size_t)buf * REGF * REG_LD;
        double v[REGF];
        {
            const double *rowp = src + (size_t)k0 * fstride;
#pragma unroll
            for (int r = 0; r < REGF; ++r) { v[r] = (mine && r < nr) ? *rowp : 0.0; rowp += fstride; }
        }
        // branch-free over the rows (rows past the end of the slab compute on zeros and are not stored): the REGF chains
        //   shift -> difference -> [shuffle] -> fractional -> wrap -> [shuffle] -> Cartesian
        // are independent of each other, only the running sum at the end is serial
#pragma unroll
        for (int r = 0; r < REGF; ++r) v[r] = v[r] - s_com[3 * (k0 + r) + comp];      // translate(-cg), msd.py:237
#pragma unroll
        for (int hh = 0; hh < REGF / REG_ILP; ++hh) {           // REG_ILP independent chains at a time (registers)
            const int h = hh * REG_ILP;
            double dl[REG_ILP];
#pragma unroll
            for (int q = 0; q < REG_ILP; ++q) {
                const int r = h + q, k = first + k0 + r;
                const double e = v[r] - (r == 0 ? prev : v[r - 1]);
                double d;
                if (CELL == 2) {
                    const double g = np_mod1(e * (comp == 0 ? i0 : comp == 1 ? i1 : i2) - shift) + shift;      // wrap_disp_diag: the diagonal entries
                    d = g * (comp == 0 ? c0 : comp == 1 ? c1 : c2);
                } else {
                    const double ex = __shfl_sync(0xffffffffu, e, base), ey = __shfl_sync(0xffffffffu, e, base + 1),
                                 ez = __shfl_sync(0xffffffffu, e, min(base + 2, 31));
                    double j0 = i0, j1 = i1, j2 = i2, b0 = c0, b1 = c1, b2 = c2;
                    if (CELL == 0) {
                        const MsdGeom &G = geom[min(max(k - 1, 0), first + count - 1)];     // cell of frame k-1 wraps k-1 -> k (rows past the slab read a valid record too)
                        j0 = G.inv[comp]; j1 = G.inv[3 + comp]; j2 = G.inv[6 + comp];
                        b0 = G.cell[comp]; b1 = G.cell[3 + comp]; b2 = G.cell[6 + comp];
                    }
                    const double g = np_mod1(((ex * j0 + ey * j1) + ez * j2) - shift) + shift;
                    const double g0 = __shfl_sync(0xffffffffu, g, base), g1 = __shfl_sync(0xffffffffu, g, base + 1),
                                 g2 = __shfl_sync(0xffffffffu, g, min(base + 2, 31));
                    d = (g0 * b0 + g1 * b1) + g2 * b2;
                }
                dl[q] = k > 0 ? d : 0.0;                        // delta_0 = 0: the running sum is taken relative to the first frame
            }
#pragma unroll
            for (int q = 0; q < REG_ILP; ++q)
                if (h + q < nr) {
                    run += dl[q];
                    if (mine) tile[(h + q) * REG_LD + col] = run;
                }
        }
#pragma unroll
        for (int r = 0; r < REGF; ++r)
            if (r == nr - 1) prev = v[r];
        __syncthreads();        // the tile is complete; the other tile's write-out (previous round) was finished by everyone who got here
        {
            const int f = lane & 15, half = lane >> 4;
            if (f < nr) {
                double *dst = P + (size_t)a0 * 3 * Tp + (size_t)(first + k0 + f);
                for (int cc = 2 * warp + half; cc < ncol; cc += 2 * (REG_THREADS / 32)) dst[(size_t)cc * Tp] = tile[f * REG_LD + cc];
            }
        }
        // no second barrier: the next round fills the OTHER tile, and the round after that passes the barrier above first
    }
    if (mine) { carry[cat] = prev; carry[cat + 3] = run; }
}

