"""GPU parity of the MSD analysis against the CPU oracle, through the C ABI (1e-12 relative, north_star)."""
import numpy as np
import pytest

from conftest import random_box
from oracle import c_oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _walk(seed, T, n, tri=True, size=12.0, step=0.4, wrapped=True, nspec=3):
    rng = np.random.default_rng(seed)
    pos0, cell0, spec = random_box(seed, n, nspec, tri, size)
    pos = pos0[None] + np.cumsum(rng.normal(scale=step, size=(T, n, 3)), axis=0)
    cells = np.array([cell0 * (1.0 + 0.001 * np.sin(k)) for k in range(T)])
    if wrapped:
        for k in range(T):
            f = pos[k] @ np.linalg.inv(cells[k])
            pos[k] = (f - np.floor(f)) @ cells[k]
    masses = np.array([1.008, 12.011, 65.38, 14.007])[spec]
    return pos, cells, spec, masses


def _gpu_window(backend, pos, cells, spec, masses, S, window, unwrap, raw_sums=False):
    T = len(pos)
    with backend.msd_open(T, masses, spec, S, cells) as s:
        s.load(0, pos[:T // 2])
        s.load(T // 2, pos[T // 2:])
        new_pos = None
        if unwrap:
            s.unwrap()
            new_pos = s.get_positions()
        sums = s.com_sums()
        com = sums[:, :3] / sums[:, 3:4]
        s.set_com(com)
        raw = s.window(np.asarray(window, dtype=np.int32))
    if raw_sums:
        return raw, com, new_pos
    n_of = np.bincount(spec, minlength=S).astype(np.float64)
    msd = raw / n_of[:, None] / (T - np.asarray(window, dtype=np.float64))[None, :]
    return msd, com, new_pos


@pytest.mark.parametrize("seed,T,n,tri,unwrap", [
    (1, 40, 50, False, False),
    (2, 65, 33, True, False),
    (3, 50, 70, True, True),
    (4, 97, 1, True, False),
    (5, 33, 300, False, True),
])
def test_window_msd(backend, seed, T, n, tri, unwrap):
    S = 3
    pos, cells, spec, masses = _walk(seed, T, n, tri)
    window = np.arange(0, T // 2, 3)
    got, com, new_pos = _gpu_window(backend, pos, cells, spec, masses, S, window, unwrap)
    want, mutated = orc.msd_window(pos, cells, masses, spec, S, window, unwrap=unwrap)
    present = np.bincount(spec, minlength=S) > 0
    assert np.all(got[present][:, 0] == 0.0)
    np.testing.assert_allclose(got[present], want[present], rtol=RTOL, atol=1e-13)
    if unwrap:
        # oracle returns unwrapped AND com-shifted positions; undo the shift with the GPU's own com
        np.testing.assert_allclose(new_pos - com[:, None, :], mutated, rtol=0, atol=1e-9)


@pytest.mark.parametrize("T,delta,env", [
    (200, 1, {}),                                                  # 100 window lengths: several groups and passes
    (203, 7, {}),                                                  # remainder frames after the last full super-row
    (161, 5, {"AMOFB_MSD_AP_NWT": "5", "AMOFB_MSD_AP_NG": "2", "AMOFB_MSD_AP_WPG": "3"}),   # 16 windows in 2 passes of 2 x 5
    (161, 5, {"AMOFB_MSD_AP_NWT": "13", "AMOFB_MSD_AP_NG": "1", "AMOFB_MSD_AP_WPG": "16"}),
    (161, 5, {"AMOFB_MSD_AP_NWT": "9", "AMOFB_MSD_AP_NG": "4", "AMOFB_MSD_AP_WPG": "1"}),   # a group with no requested window in pass 1
    (120, 10, {"AMOFB_MSD_NO_AP": "1"}),                           # generic kernel on the same kind of request
])
def test_window_msd_arithmetic_progression(backend, monkeypatch, T, delta, env):
    """WindowMsd's own request (window = arange(0, T//2, delta), msd.py:176-178) through every launch shape of the
    register-tiled kernel, against the oracle."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    S = 3
    pos, cells, spec, masses = _walk(40 + T + delta, T, 37, True)
    window = np.arange(0, T // 2, delta)
    got, _, _ = _gpu_window(backend, pos, cells, spec, masses, S, window, False)
    want, _ = orc.msd_window(pos, cells, masses, spec, S, window)
    assert np.all(got[:, 0] == 0.0)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-13)


@pytest.mark.parametrize("tri,unwrap", [(False, False), (False, True), (True, False)])
def test_window_msd_fixed_cell(backend, tri, unwrap):
    """NVT trajectories: one cell for every frame (geometry kept in registers) and, for an orthorhombic box along the
    axes, the diagonal form of the wrap -- same numbers as the general P8 arithmetic of the oracle."""
    S = 3
    T = 150
    pos, cells, spec, masses = _walk(300 + int(tri), T, 53, tri, wrapped=True)
    cells[:] = cells[0]
    f = np.einsum('kni,ij->knj', pos, np.linalg.inv(cells[0]))
    pos = np.einsum('kni,ij->knj', f - np.floor(f), cells[0])
    window = np.arange(0, T // 2, 4)
    got, com, new_pos = _gpu_window(backend, pos, cells, spec, masses, S, window, unwrap)
    want, mutated = orc.msd_window(pos, cells, masses, spec, S, window, unwrap=unwrap)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-13)
    if unwrap:
        np.testing.assert_allclose(new_pos - com[:, None, :], mutated, rtol=0, atol=1e-9)


def test_window_msd_long_series_properties(backend, monkeypatch):
    """A C5-shaped request (unwrapped random walk, fixed orthorhombic box, window = arange(0, T//2, 50)) at a size the
    oracle does not finish in seconds: the tiled kernel must agree with the generic one to 1e-12 and follow the
    random-walk law MSD(m) = 3 sigma^2 m (T-m-1)/(T-m) (quirk Q4) within the sampling noise."""
    T, n, S, sigma = 1200, 3000, 4, 0.05
    rng = np.random.default_rng(5)
    pos = 100.0 + np.cumsum(rng.normal(scale=sigma, size=(T, n, 3)), axis=0)
    cells = np.broadcast_to(np.diag([231.0, 246.0, 277.0]), (T, 3, 3)).copy()
    spec = (np.arange(n) % S).astype(np.uint8)
    masses = np.array([1.008, 12.011, 14.007, 65.38])[spec]
    window = np.arange(0, T // 2, 50)
    tiled, _, _ = _gpu_window(backend, pos, cells, spec, masses, S, window, False)
    monkeypatch.setenv("AMOFB_MSD_NO_AP", "1")
    generic, _, _ = _gpu_window(backend, pos, cells, spec, masses, S, window, False)
    np.testing.assert_allclose(tiled, generic, rtol=RTOL, atol=1e-13)
    m = window.astype(np.float64)
    law = 3.0 * sigma ** 2 * m * (T - m - 1) / (T - m)
    assert np.all(tiled[:, 0] == 0.0)
    np.testing.assert_allclose(tiled[:, 1:].mean(axis=0), law[1:], rtol=0.05)      # COM removal costs O(1/n)


def test_window_msd_irregular_windows(backend):
    """window lengths that are NOT an arithmetic progression (only reachable through the C ABI) -> generic kernel;
    lengths >= T contribute nothing"""
    S = 3
    pos, cells, spec, masses = _walk(77, 90, 41, True)
    window = np.array([0, 2, 5, 11, 12, 40, 89, 90, 500])
    raw, _, _ = _gpu_window(backend, pos, cells, spec, masses, S, window, False, raw_sums=True)
    ok = window < 90
    want, _ = orc.msd_window(pos, cells, masses, spec, S, window[ok])
    n_of = np.bincount(spec, minlength=S).astype(np.float64)
    got = raw[:, ok] / n_of[:, None] / (90 - window[ok].astype(np.float64))[None, :]
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-13)
    assert np.all(raw[:, ~ok] == 0.0)


def test_msd_get_positions_roundtrip(backend):
    pos, cells, spec, masses = _walk(9, 37, 45, True)
    with backend.msd_open(len(pos), masses, spec, 3, cells) as s:
        s.load(0, pos)
        back = s.get_positions()
        sums = s.com_sums()
    assert np.array_equal(back, pos)
    want = (masses[None, :, None] * pos).sum(axis=1)
    np.testing.assert_allclose(sums[:, :3], want, rtol=1e-13)
    np.testing.assert_allclose(sums[:, 3], masses.sum(), rtol=1e-15)


def test_long_series_global_path(backend, monkeypatch):
    """T too long for shared memory staging -> the global-memory variant of the window kernel."""
    monkeypatch.setenv("AMOFB_MSD_NO_SMEM", "1")
    pos, cells, spec, masses = _walk(12, 60, 40, True)
    window = np.arange(0, 30, 5)
    got, _, _ = _gpu_window(backend, pos, cells, spec, masses, 3, window, False)
    want, _ = orc.msd_window(pos, cells, masses, spec, 3, window)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-13)


def test_direct_msd(backend):
    pos, cells, spec, masses = _walk(21, 45, 60, tri=False, step=0.3)
    with backend.msd_open(len(pos), masses, spec, 3, cells) as s:
        s.load(0, pos)
        raw = s.direct()
    n_of = np.bincount(spec, minlength=3)
    for sp in range(3):
        want = orc.msd_direct(pos, cells, spec, sp)
        np.testing.assert_allclose(raw[sp] / n_of[sp], want, rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(raw.sum(axis=0) / len(spec), orc.msd_direct(pos, cells, spec, -1), rtol=RTOL, atol=1e-13)
    tri = cells.copy()
    tri[:, 1, 0] = 0.5
    with backend.msd_open(len(pos), masses, spec, 3, tri) as s:
        s.load(0, pos)
        with pytest.raises(ValueError):
            s.direct()


def test_state_machine_errors(backend):
    pos, cells, spec, masses = _walk(30, 20, 10, False)
    with backend.msd_open(len(pos), masses, spec, 3, cells) as s:
        s.load(0, pos)
        with pytest.raises(RuntimeError):
            s.window(np.array([0, 1], dtype=np.int32))          # no centre of mass yet
        with pytest.raises(ValueError):
            s.load(15, pos[:10])                                  # outside [0, T)


# ---- streaming path (amofb_msd_slab_*): mass sums on the way in, shift + wrap + running sum fused into the transposition,
# ---- autocorrelation window kernel --------------------------------------------------------------------------------
def _gpu_stream(backend, pos, cells, spec, masses, S, window, slab, device=False):
    T = len(pos)
    com = np.empty((T, 3))
    with backend.msd_open(T, masses, spec, S, cells) as s:
        assert s.slab_frames() >= 1
        for a in range(0, T, slab):
            b = min(T, a + slab)
            if device:
                d = backend.ctx.device_alloc(pos[a:b].nbytes)
                backend.ctx.h2d(d, pos[a:b])
                sums = s.slab_sums(a, (d.value, b - a))
            else:
                sums = s.slab_sums(a, pos[a:b])
            com[a:b] = sums[:, :3] / sums[:, 3:4]
            s.slab_commit(com[a:b])
            if device:
                backend.ctx.sync()
                backend.ctx.device_free(d)
        raw = s.window(np.asarray(window, dtype=np.int32))
    n_of = np.bincount(spec, minlength=S).astype(np.float64)
    return raw / n_of[:, None] / (T - np.asarray(window, dtype=np.float64))[None, :], com


@pytest.mark.parametrize("seed,T,n,tri,fixed,slab,delta,device", [
    (1, 80, 50, False, False, 80, 3, False),       # one slab, a cell per frame
    (2, 131, 33, True, False, 17, 5, False),       # odd T, slabs that are not multiples of the 32-frame rounds, carry across slabs
    (3, 150, 70, True, True, 64, 4, True),         # one triclinic cell for every frame, device-resident slabs
    (4, 97, 1, True, False, 10, 2, False),         # a single atom
    (5, 200, 200, False, True, 33, 1, False),      # orthorhombic box along the axes (diagonal wrap), 100 window lengths
    (6, 90, 129, True, False, 90, 7, True),        # atoms not a multiple of the 64-atom blocks
])
def test_streaming_path_matches_oracle(backend, seed, T, n, tri, fixed, slab, delta, device):
    S = 3
    pos, cells, spec, masses = _walk(500 + seed, T, n, tri)
    if fixed:
        cells[:] = cells[0]
        f = np.einsum('kni,ij->knj', pos, np.linalg.inv(cells[0]))
        pos = np.einsum('kni,ij->knj', f - np.floor(f), cells[0])
    window = np.arange(0, T // 2, delta)
    got, com = _gpu_stream(backend, pos, cells, spec, masses, S, window, slab, device)
    want, _ = orc.msd_window(pos, cells, masses, spec, S, window)
    present = np.bincount(spec, minlength=S) > 0
    assert np.all(got[present][:, 0] == 0.0)
    np.testing.assert_allclose(got[present], want[present], rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(com, (masses[None, :, None] * pos).sum(axis=1) / masses.sum(), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("T,n,delta,kb,nwt", [
    (1201, 37, 25, 6, 0),        # odd T, 24 window lengths -> 25 sums per thread
    (1200, 20, 290, 8, 0),       # a stride beyond the block size: 3 window lengths
    (1000, 64, 20, 10, 13),      # 25 window lengths forced into two passes of 13
    (640, 30, 4, 8, 0),          # 80 window lengths: three passes of 32, later passes start at a window offset
    (333, 9, 1, 10, 32),         # stride 1, 166 window lengths
])
def test_wide_window_kernel_shapes(backend, monkeypatch, T, n, delta, kb, nwt):
    """k_msd_window_wide (two series buffers, all windows of a pass in one register tile) against the oracle AND against the
    narrow kernel it replaced, for every tile height and for requests of one and of several passes."""
    S = 2
    pos, cells, spec, masses = _walk(900 + T, T, n, True, nspec=S)
    window = np.arange(0, T // 2, delta)
    monkeypatch.setenv("AMOFB_MSD_WIDE_KB", str(kb))
    if nwt:
        monkeypatch.setenv("AMOFB_MSD_WIDE_NWT", str(nwt))
    wide, _ = _gpu_stream(backend, pos, cells, spec, masses, S, window, 96)
    monkeypatch.setenv("AMOFB_MSD_NO_WIDE", "1")
    narrow, _ = _gpu_stream(backend, pos, cells, spec, masses, S, window, 96)
    want, _ = orc.msd_window(pos, cells, masses, spec, S, window)
    np.testing.assert_allclose(wide, want, rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(narrow, want, rtol=RTOL, atol=1e-13)


def test_streaming_forms_agree_and_fall_back(backend, monkeypatch):
    """The autocorrelation form must agree with the difference form, and a request whose windows are small against the
    squares they are taken from (ballistic drift, lag 1) must come out right all the same (the library re-runs it in the
    difference form)."""
    S, T, n = 2, 400, 40
    pos, cells, spec, masses = _walk(71, T, n, True, size=40.0, step=0.05, wrapped=False, nspec=S)
    window = np.arange(0, T // 2, 10)
    dot, _ = _gpu_stream(backend, pos, cells, spec, masses, S, window, 128)
    monkeypatch.setenv("AMOFB_MSD_NO_DOT", "1")
    diff, _ = _gpu_stream(backend, pos, cells, spec, masses, S, window, 128)
    monkeypatch.delenv("AMOFB_MSD_NO_DOT")
    want, _ = orc.msd_window(pos, cells, masses, spec, S, window)
    np.testing.assert_allclose(dot, want, rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(diff, want, rtol=RTOL, atol=1e-13)
    # two species moving against each other at constant velocity (the centre of mass stays put), tiny jitter:
    # |R_k|^2 grows like k^2 while MSD(1) stays v^2
    rng = np.random.default_rng(3)
    vel = np.where(spec[:, None] == 0, 1.0, -1.0) * np.array([0.01, 0.02, -0.015]) / masses[:, None]
    pos = pos[0][None] + np.arange(T)[:, None, None] * vel[None] + rng.normal(scale=1e-4, size=(T, n, 3))
    window = np.arange(0, T // 2, 1)
    got, _ = _gpu_stream(backend, pos, cells, spec, masses, S, window, 256)
    want, _ = orc.msd_window(pos, cells, masses, spec, S, window)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-15)


def test_streaming_state_machine(backend):
    pos, cells, spec, masses = _walk(31, 40, 12, False)
    with backend.msd_open(len(pos), masses, spec, 3, cells) as s:
        with pytest.raises(ValueError):
            s.slab_sums(5, pos[5:10])                             # slabs come in frame order
        sums = s.slab_sums(0, pos[:16])
        with pytest.raises(RuntimeError):
            s.load(0, pos)                                        # the two paths do not mix
        s.slab_sums_begin(16, pos[16:30])                         # two slabs may await their commit ...
        with pytest.raises(RuntimeError):
            s.slab_sums_begin(30, pos[30:])                       # ... a third may not
        s.slab_commit(sums[:, :3] / sums[:, 3:4])
        with pytest.raises(RuntimeError):
            s.window(np.array([0, 2, 4], dtype=np.int32))         # 24 frames are still missing
        with pytest.raises(RuntimeError):
            s.slab_commit(sums[:14, :3])                          # the sums of the oldest slab were never fetched
        sums = s.slab_sums_wait(14)
        s.slab_commit(sums[:, :3] / sums[:, 3:4])
        sums = s.slab_sums(30, pos[30:])
        s.slab_commit(sums[:, :3] / sums[:, 3:4])
        with pytest.raises(ValueError):
            s.window(np.array([0, 2, 5], dtype=np.int32))         # not an arithmetic progression: the other path's job
        raw = s.window(np.array([0, 2, 4, 6], dtype=np.int32))
    assert raw.shape == (3, 4) and np.all(raw[:, 0] == 0.0)
