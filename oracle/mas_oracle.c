/*
 * oracle/mas_oracle.c -- CPU restatement of the reference's Monotonic
 * Alignment Search.  TEST INFRASTRUCTURE ONLY: the product
 * (face-gan-tts_b200/) never links, loads or calls this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg do, as the checker.
 *
 * Parity status: PINNED.  The reference ships no golden vectors (SURVEY.md
 * section 4), so this restatement is pinned against the reference's own
 * compiled core.pyx (oracle/build_ref.py -> oracle/_ref/) in
 * tests/test_oracle.py and against the fixtures in tests/golden/ that the
 * compiled reference generated (tests/golden/make_golden.py).
 *
 * Follows /root/reference/model/monotonic_align/core.pyx line by line:
 *   mas_oracle_each        <- maximum_path_each   core.pyx:9-35
 *   mas_oracle_batch       <- maximum_path_c      core.pyx:40-45
 * Differences (deliberate, none changes a defined result):
 *   - none in the arithmetic: like the reference, `values` IS clobbered
 *     (accumulated in place, core.pyx:30) and `paths` must come pre-zeroed.
 *   - t_x > t_y (undefined behaviour in the reference: core.pyx:34 reads
 *     value[index,-1] with wraparound off) returns 1 for that item and
 *     leaves its path untouched instead of reading out of bounds.
 *
 * Build: gcc -O2 -shared -fPIC -o libmas_oracle.so mas_oracle.c   (oracle/Makefile)
 */
#include <stddef.h>

/* Cython lowers `max(v_cur, v_prev)` to (v_prev > v_cur) ? v_prev : v_cur
 * (checked in the generated C, oracle/_ref/gen/core.c): NaN in either operand
 * selects v_cur. */
static float ref_max(float v_cur, float v_prev) { return (v_prev > v_cur) ? v_prev : v_cur; }

/* core.pyx:9-35.  path: [Tx,Ty] int32 pre-zeroed (row stride Ty); value:
 * [Tx,Ty] fp32, accumulated in place. */
int mas_oracle_each(int *path, float *value, int Ty, int t_x, int t_y, float max_neg_val)
{
    int x, y, index = t_x - 1;
    float v_prev, v_cur;
    if (t_x < 1 || t_y < 1 || t_x > t_y) return 1;

    for (y = 0; y < t_y; ++y) {                                 /* core.pyx:17 */
        int lo = t_x + y - t_y; if (lo < 0) lo = 0;             /* core.pyx:18 */
        int hi = (y + 1 < t_x) ? y + 1 : t_x;
        for (x = lo; x < hi; ++x) {
            if (x == y) v_cur = max_neg_val;                    /* core.pyx:19-22 */
            else        v_cur = value[(size_t)x * Ty + (y - 1)];
            if (x == 0) v_prev = (y == 0) ? 0.f : max_neg_val;  /* core.pyx:23-29 */
            else        v_prev = value[(size_t)(x - 1) * Ty + (y - 1)];
            value[(size_t)x * Ty + y] = ref_max(v_cur, v_prev) + value[(size_t)x * Ty + y]; /* :30 */
        }
    }
    for (y = t_y - 1; y >= 0; --y) {                            /* core.pyx:32 */
        path[(size_t)index * Ty + y] = 1;                       /* core.pyx:33 */
        if (index != 0 && (index == y ||
             value[(size_t)index * Ty + (y - 1)] < value[(size_t)(index - 1) * Ty + (y - 1)]))
            index = index - 1;                                  /* core.pyx:34-35 */
    }
    return 0;
}

/* core.pyx:40-45 (serial, as the reference is actually built: no -fopenmp). */
int mas_oracle_batch(int *paths, float *values, const int *t_xs, const int *t_ys,
                     int B, int Tx, int Ty, float max_neg_val)
{
    int i, bad = 0;
    for (i = 0; i < B; ++i)
        bad += mas_oracle_each(paths + (size_t)i * Tx * Ty, values + (size_t)i * Tx * Ty,
                               Ty, t_xs[i], t_ys[i], max_neg_val);
    return bad;
}

/* Helpers the tests use to compare compact outputs with the dense path.
 * durations[x] = sum_y path[x,y]  (what face_tts.py:176 recovers with a dense
 * re-read); frame_token[y] = the x with path[x,y]==1, or -1. */
void mas_oracle_durations(const int *path, int Tx, int Ty, int *dur, int *frame_token)
{
    int x, y;
    for (x = 0; x < Tx; ++x) dur[x] = 0;
    for (y = 0; y < Ty; ++y) frame_token[y] = -1;
    for (x = 0; x < Tx; ++x)
        for (y = 0; y < Ty; ++y)
            if (path[(size_t)x * Ty + y]) { dur[x] += 1; frame_token[y] = x; }
}
