// Host-side text ingest of XYZ-family trajectories (include/amofb.h: amofb_xyz_parse).  Replaces the per-frame Python parsing
// of ase.io.read(..., format='xyz') behind Trajectory.from_traj / read_lammps_traj / read_cp2k_traj
// (/root/reference/amof/trajectory.py:48-60,193-228).  Plain C++: no device work, no CUDA types.
#include <atomic>
#include <charconv>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/amofb.h"

namespace {

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r'; }

inline const char *skip_blank(const char *p, const char *end) {
    while (p < end && is_blank(*p)) ++p;
    return p;
}
inline const char *skip_token(const char *p, const char *end) {
    while (p < end && !is_blank(*p) && *p != '\n') ++p;
    return p;
}
inline const char *next_line(const char *p, const char *end) {
    const void *q = memchr(p, '\n', (size_t)(end - p));
    return q ? (const char *)q + 1 : end;
}

// one decimal string -> binary64, correctly rounded (std::from_chars); Fortran exponents (1.0D+00) go through a copy
inline bool parse_double(const char *&p, const char *end, double &out) {
    const char *q = p;
    if (q < end && *q == '+') ++q;
    auto r = std::from_chars(q, end, out, std::chars_format::general);
    if (r.ec != std::errc()) return false;
    if (r.ptr < end && (*r.ptr == 'D' || *r.ptr == 'd')) {
        char tmp[64];
        const char *t = skip_token(q, end);
        size_t n = (size_t)(t - q);
        if (n >= sizeof tmp) return false;
        memcpy(tmp, q, n);
        tmp[n] = 0;
        for (size_t i = 0; i < n; ++i)
            if (tmp[i] == 'D' || tmp[i] == 'd') tmp[i] = 'e';
        auto r2 = std::from_chars(tmp, tmp + n, out, std::chars_format::general);
        if (r2.ec != std::errc() || r2.ptr != tmp + n) return false;
        p = t;
        return true;
    }
    p = r.ptr;
    return p == end || is_blank(*p) || *p == '\n';
}

// returns 0, or 1 (malformed) / 2 (atom order changed)
int parse_frame(const char *p, const char *end, int n_atoms, int pos_col, char *symbols, bool fill_symbols, double *out) {
    p = next_line(p, end);          // count line
    p = next_line(p, end);          // comment line
    for (int a = 0; a < n_atoms; ++a) {
        p = skip_blank(p, end);
        if (p >= end || *p == '\n') return 1;
        const char *t = skip_token(p, end);
        size_t n = (size_t)(t - p);
        char *sym = symbols + 8 * (size_t)a;
        if (fill_symbols) {
            if (n > 7) return 1;
            memset(sym, 0, 8);
            memcpy(sym, p, n);
        } else if (n > 7 || memcmp(sym, p, n) != 0 || sym[n] != 0) {
            return 2;
        }
        p = t;
        for (int c = 1; c < pos_col; ++c) {
            p = skip_blank(p, end);
            if (p >= end || *p == '\n') return 1;
            p = skip_token(p, end);
        }
        for (int c = 0; c < 3; ++c) {
            p = skip_blank(p, end);
            if (p >= end || *p == '\n') return 1;
            if (!parse_double(p, end, out[3 * (size_t)a + c])) return 1;
        }
        p = next_line(p, end);
    }
    return 0;
}

}  // namespace

extern "C" int amofb_xyz_parse(const char *text, const int64_t *frame_off, int n_frames, int n_atoms, int pos_col, char *symbols,
                               int symbols_known, double *positions, int threads, int *bad_frame) {
    if (bad_frame) *bad_frame = -1;
    if (!text || !frame_off || !symbols || !positions || n_frames < 0 || n_atoms <= 0 || pos_col < 1) return AMOFB_ERR_ARG;
    if (n_frames == 0) return AMOFB_OK;
    const size_t fstride = 3 * (size_t)n_atoms;
    int first = 0;
    if (!symbols_known) {           // the first frame names the atoms; the others are checked against it
        int rc = parse_frame(text + frame_off[0], text + frame_off[1], n_atoms, pos_col, symbols, true, positions);
        if (rc) {
            if (bad_frame) *bad_frame = 0;
            return AMOFB_ERR_ARG;
        }
        first = 1;
    }
    if (threads <= 0) {
        unsigned hc = std::thread::hardware_concurrency();
        threads = (int)(hc ? (hc < 16 ? hc : 16) : 1);
    }
    if (threads > n_frames - first) threads = n_frames - first > 0 ? n_frames - first : 1;
    std::atomic<int> next(first), bad(-1);
    auto work = [&]() {
        for (;;) {
            int f = next.fetch_add(1, std::memory_order_relaxed);
            if (f >= n_frames || bad.load(std::memory_order_relaxed) >= 0) return;
            int rc = parse_frame(text + frame_off[f], text + frame_off[f + 1], n_atoms, pos_col, symbols, false, positions + fstride * (size_t)f);
            if (rc) {
                int expect = -1;
                bad.compare_exchange_strong(expect, f);
                return;
            }
        }
    };
    if (threads == 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work);
        for (auto &t : pool) t.join();
    }
    if (bad.load() >= 0) {
        if (bad_frame) *bad_frame = bad.load();
        return AMOFB_ERR_ARG;
    }
    return AMOFB_OK;
}

namespace {

// newlines in [p, end)
int64_t count_newlines(const char *p, const char *end) {
    int64_t n = 0;
    while (p < end) {
        const void *q = memchr(p, '\n', (size_t)(end - p));
        if (!q) break;
        p = (const char *)q + 1;
        ++n;
    }
    return n;
}

// frame starts after the newlines of [p, end): `until` newlines are left before the next frame starts
void emit_starts(const char *text, const char *p, const char *end, int64_t until, int64_t period, int64_t base, std::vector<int64_t> &out) {
    while (p < end) {
        const void *q = memchr(p, '\n', (size_t)(end - p));
        if (!q) break;
        p = (const char *)q + 1;
        if (--until == 0) {
            out.push_back(base + (p - text));
            until = period;
        }
    }
}

}  // namespace

extern "C" int amofb_xyz_index(const char *text, int64_t len, int64_t lines_before, int64_t period, int64_t base, int64_t *starts,
                               int64_t capacity, int64_t *n_starts, int64_t *n_lines) {
    if (!text || len < 0 || period <= 0 || !n_starts || !n_lines || (capacity > 0 && !starts)) return AMOFB_ERR_ARG;
    // large blocks: the host cores count the newlines of a slice each, then -- knowing how many lines precede their slice --
    // emit the frame starts of it (two memchr passes in parallel instead of one serial one)
    unsigned hc = std::thread::hardware_concurrency();
    int threads = (int)(hc ? (hc < 16 ? hc : 16) : 1);
    if (len < (int64_t)(4 << 20)) threads = 1;
    std::vector<int64_t> count((size_t)threads, 0);
    std::vector<std::vector<int64_t>> found((size_t)threads);
    auto lo = [&](int t) { return text + len * t / threads; };
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; ++t) pool.emplace_back([&, t]() { count[(size_t)t] = count_newlines(lo(t), lo(t + 1)); });
        count[0] = count_newlines(lo(0), lo(1));
        for (auto &th : pool) th.join();
    }
    {
        std::vector<int64_t> before((size_t)threads + 1, lines_before);
        for (int t = 0; t < threads; ++t) before[(size_t)t + 1] = before[(size_t)t] + count[(size_t)t];
        std::vector<std::thread> pool;
        auto run = [&](int t) { emit_starts(text, lo(t), lo(t + 1), period - before[(size_t)t] % period, period, base, found[(size_t)t]); };
        for (int t = 1; t < threads; ++t) pool.emplace_back(run, t);
        run(0);
        for (auto &th : pool) th.join();
        *n_lines = before[(size_t)threads] - lines_before;
    }
    int64_t n = 0;
    for (auto &v : found)
        for (int64_t x : v) {
            if (n < capacity) starts[n] = x;
            ++n;
        }
    *n_starts = n;
    return n > capacity ? AMOFB_ERR_MEMORY : AMOFB_OK;
}
