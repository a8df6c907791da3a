"""
A stand-in with the interface of ``amof_b200._lib.GpuBackend`` that answers from the CPU oracle.

TEST INFRASTRUCTURE: it lets the ``-m "not gpu"`` suite exercise the host-side logic of the analysis classes
(argument handling, quirks Q1-Q7, DataFrame schemas, file round-trips, rank sharding and reductions) on a machine
without a GPU.  It is installed with ``amof_b200._lib._set_backend_for_tests`` by the tests only; the package itself
never imports it, and the GPU parity tests (``-m gpu``) never use it.
"""
import numpy as np

from oracle import c_oracle as orc


class OracleBackend:
    name = "oracle (tests only)"
    ctx = None

    def pair_counts(self, species, n_species, chunks, rmax=0.0, nbins=0, cn_cutoff=None):
        S = int(n_species)
        hist = np.zeros((S, S, int(nbins)), dtype=np.uint64) if nbins > 0 else None
        cn, nf, vs = [], 0, 0.0
        for pos, cell in chunks:
            pos = np.asarray(pos, dtype=np.float64)
            cell = np.asarray(cell, dtype=np.float64).reshape(-1, 3, 3)
            for f in range(len(cell)):
                if hist is not None:
                    hist += orc.rdf_hist(pos[f], cell[f], species, S, rmax, nbins)
                if cn_cutoff is not None:
                    cn.append(orc.cn_counts(pos[f], cell[f], species, S, cn_cutoff))
                vs += abs(np.linalg.det(cell[f]))
                nf += 1
        cn = (np.array(cn, dtype=np.uint64).reshape(nf, S, S) if cn_cutoff is not None else None)
        return {"hist": hist, "cn": cn, "n_frames": nf, "volume_sum": vs}

    def pair_counts_each(self, species, n_species, chunks, rmax, nbins):
        for chunk in chunks:
            yield self.pair_counts(species, n_species, [chunk], rmax=rmax, nbins=nbins)

    def bad_counts(self, species, n_species, chunks, cutoff, triples, dtheta, nbins):
        hist = np.zeros((len(triples), 33, int(nbins)), dtype=np.uint64)
        dropped = np.zeros(len(triples), dtype=np.uint64)
        nf = 0
        for pos, cell in chunks:
            pos = np.asarray(pos, dtype=np.float64)
            cell = np.asarray(cell, dtype=np.float64).reshape(-1, 3, 3)
            for f in range(len(cell)):
                for t, (A, B) in enumerate(triples):
                    _, d = orc.bad_hist(pos[f], cell[f], species, n_species, cutoff, A, B, dtheta, nbins, hist=hist[t])
                    dropped[t] += d
                nf += 1
        return hist, dropped, nf

    def neighbour_list(self, species, n_species, positions, cell, cutoff, quantities=False):
        species = np.asarray(species, dtype=np.uint8)
        i, j, d, S = orc.neighbour_pairs(positions, cell, species, int(n_species), cutoff, quantities=True)
        offsets = np.zeros(len(species) + 1, dtype=np.int64)
        np.cumsum(np.bincount(i, minlength=len(species)), out=offsets[1:])
        if quantities:
            return offsets, j.astype(np.int32), d, S
        return offsets, j.astype(np.int32)

    def msd_open(self, n_frames, masses, species, n_species, cells):
        return _Session(int(n_frames), np.asarray(masses, dtype=np.float64), np.asarray(species, dtype=np.uint8),
                        int(n_species), np.asarray(cells, dtype=np.float64).reshape(int(n_frames), 3, 3))


class _Session:
    def __init__(self, T, masses, species, S, cells):
        self.T, self.masses, self.species, self.S, self.cells = T, masses, species, S, cells
        self.pos = np.zeros((T, len(species), 3))
        self.com = None

    def load(self, first, pos):
        pos = np.asarray(pos, dtype=np.float64)
        self.pos[first:first + len(pos)] = pos

    def unwrap(self):
        d = orc.delta_pos(self.pos, self.cells)
        self.pos = np.cumsum(d, axis=0)

    def com_sums(self):
        out = np.zeros((self.T, 4))
        out[:, :3] = (self.masses[None, :, None] * self.pos).sum(axis=1)
        out[:, 3] = self.masses.sum()
        return out

    def set_com(self, com):
        self.com = np.asarray(com, dtype=np.float64).reshape(self.T, 3)

    def window(self, window):
        q = self.pos - self.com[:, None, :]
        R = np.cumsum(orc.delta_pos(q, self.cells), axis=0)
        out = np.zeros((self.S, len(window)))
        for w, m in enumerate(window):
            m = int(m)
            if m < self.T - 1:
                d = R[m + 1:] - R[1:self.T - m]
                sq = (d * d).sum(axis=2).sum(axis=0)
                for s in range(self.S):
                    out[s, w] = sq[self.species == s].sum()
        return out

    def direct(self):
        out = np.zeros((self.S, self.T))
        for s in range(self.S):
            n = int((self.species == s).sum())
            if n:
                out[s] = orc.msd_direct(self.pos, self.cells, self.species, s) * n
        return out

    def get_positions(self):
        return self.pos.copy()

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
