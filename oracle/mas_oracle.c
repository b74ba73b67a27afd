/* TEST INFRASTRUCTURE ONLY -- CPU oracle of the monotonic alignment search.  Imported by tests/, __graft_entry__.smoke()
 * and bench.py's CPU legs; never by the product path (emojivoice_b200/ fails loudly without its CUDA library).
 *
 * Plain-C restatement of Matcha-TTS/matcha/utils/monotonic_align/core.pyx:
 *   maximum_path_each  core.pyx:11-38   (forward DP :22-34, backward trace :36-39)
 *   maximum_path_c     core.pyx:42-47   (loop over the batch)
 * Pinned against the reference itself: oracle/build_oracle.py cythonizes the unmodified core.pyx into oracle/_ref/ in
 * the build container, tests/test_mas.py compares both on random inputs, and tests/golden/mas_*.npz holds paths the
 * reference produced (scripts/make_golden_mas.py).
 *
 * Arithmetic notes (they decide bit-exactness):
 *   - Cython compiles max(v_cur, v_prev) to  (v_prev > v_cur) ? v_prev : v_cur   (second argument wins only if greater);
 *   - value[x, y] = max(...) + value[x, y] is ONE float32 addition per cell;
 *   - boundscheck/wraparound are off in the reference: t_y >= t_x is required (otherwise it reads out of bounds).
 */
#include <stdint.h>

static void maximum_path_each(int32_t* path, float* value, int t_x, int t_y, int ld, float max_neg_val) {
  int index = t_x - 1;
  for (int y = 0; y < t_y; ++y) {
    int lo = t_x + y - t_y; if (lo < 0) lo = 0;
    int hi = y + 1; if (hi > t_x) hi = t_x;
    for (int x = lo; x < hi; ++x) {
      float v_cur, v_prev;
      if (x == y) v_cur = max_neg_val; else v_cur = value[x * ld + y - 1];
      if (x == 0) v_prev = (y == 0) ? 0.0f : max_neg_val; else v_prev = value[(x - 1) * ld + y - 1];
      const float m = (v_prev > v_cur) ? v_prev : v_cur;
      value[x * ld + y] = m + value[x * ld + y];
    }
  }
  for (int y = t_y - 1; y >= 0; --y) {
    path[index * ld + y] = 1;
    if (index != 0 && (index == y || value[index * ld + y - 1] < value[(index - 1) * ld + y - 1])) index = index - 1;
  }
}

/* paths (b, Tx, Ty) int32 zero-initialised by the caller; values (b, Tx, Ty) float32, modified in place like the
 * reference's private copy. */
void mas_oracle_maximum_path(int32_t* paths, float* values, const int32_t* t_xs, const int32_t* t_ys, int b, int Tx, int Ty,
                             float max_neg_val) {
  for (int i = 0; i < b; ++i)
    if (t_xs[i] > 0 && t_ys[i] > 0)
      maximum_path_each(paths + (long)i * Tx * Ty, values + (long)i * Tx * Ty, t_xs[i], t_ys[i], Ty, max_neg_val);
}
