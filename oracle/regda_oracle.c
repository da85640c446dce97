/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the integer / exact-compare
 * stages of RegDA's self-training inner step.  Never linked into, imported by or called
 * from the product (regda_b200/); used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker.
 *
 * Parity status: PINNED.  Every function here is checked against outputs of the
 * unmodified reference (imported through oracle/ref_loader.py in the build container)
 * on the committed fixtures under tests/golden/ (tests/test_oracle_golden.py), and
 * directly against the live reference when /root/reference is present
 * (tests/test_oracle_vs_reference.py).
 *
 * Functions restate:
 *   oracle_lrh              regda/utils/local_region_homog.py:107-152 (Homogenizer)
 *   oracle_pseudo_select    regda/gast/pseudo_generation.py:59-93     (pseudo_selection)
 *   oracle_downscale_label  regda/gast/alignment.py:456-481           (DownscaleLabel)
 *   oracle_class_count      regda/gast/balance.py:43-51               (ClassBalance._local_freq counts)
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 * All float32 steps are written as explicit single-precision operations so that no
 * double-rounding or contraction can creep in.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_ERR_LABEL_RANGE 1   /* reference: one_hot() raises RuntimeError        */
#define ORACLE_ERR_REGION_RANGE 2  /* reference: scatter_add_ index out of range      */
#define ORACLE_ERR_ALLOC 3
#define ORACLE_ERR_PROB_RANGE 4    /* reference: assert mask.max()<=1 and mask.min()>=0 */

/* ---------------------------------------------------------------------------------------
 * Local Region Homogenizing.  local_region_homog.py:125-152.
 *   :114-118  ignore_label -> class_num, then one_hot(class_num+1)[:, :-1]:
 *             a label must lie in [0, class_num] after that remap, class_num itself
 *             contributes to no bin.
 *   :140      per-image histogram counts[b][region][class]  (int64 scatter-add)
 *   :141-144  valid = sum_c counts; (max, argmax) with first-max tie rule;
 *             ratio = float32(max) / (float32(valid) + 1e-5f)   -- float32 arithmetic
 *             winner := ignore_label where ratio < float32(percent)
 *   :147-151  gather winner back; region 0 -> ignore; where(== ignore) keep input label.
 * --------------------------------------------------------------------------------------- */
int oracle_lrh(const int64_t *labels, const int64_t *regions, int64_t *out,
               int64_t b, int64_t hw, int class_num, int64_t ignore_label, double percent)
{
    const int64_t n = b * hw;
    int64_t rmax = -1;
    for (int64_t i = 0; i < n; ++i) {
        if (regions[i] < 0) return ORACLE_ERR_REGION_RANGE;
        if (regions[i] > rmax) rmax = regions[i];
    }
    for (int64_t i = 0; i < n; ++i) {
        int64_t l = labels[i];
        if (l == ignore_label) l = class_num;
        if (l < 0 || l > class_num) return ORACLE_ERR_LABEL_RANGE;
    }
    if (n == 0) return ORACLE_OK;
    const int64_t R = rmax + 1; /* batch-global, as in the reference (unobservable) */
    int64_t *cnt = (int64_t *)calloc((size_t)(R * class_num > 0 ? R * class_num : 1), sizeof(int64_t));
    int64_t *win = (int64_t *)malloc((size_t)(R > 0 ? R : 1) * sizeof(int64_t));
    if (!cnt || !win) { free(cnt); free(win); return ORACLE_ERR_ALLOC; }
    const float pf = (float)percent;
    for (int64_t img = 0; img < b; ++img) {
        const int64_t *lab = labels + img * hw;
        const int64_t *reg = regions + img * hw;
        int64_t *o = out + img * hw;
        memset(cnt, 0, (size_t)(R * class_num) * sizeof(int64_t));
        for (int64_t i = 0; i < hw; ++i) {
            int64_t l = lab[i];
            if (l == ignore_label) l = class_num;
            if (l < class_num) cnt[reg[i] * class_num + l] += 1;
        }
        for (int64_t r = 0; r < R; ++r) {
            int64_t valid = 0, best = INT64_MIN, arg = 0;
            for (int c = 0; c < class_num; ++c) {
                int64_t v = cnt[r * class_num + c];
                valid += v;
                if (v > best) { best = v; arg = c; } /* strict > keeps the first maximum */
            }
            if (class_num == 0) { best = 0; }
            volatile float den = (float)valid + 1e-5f;
            volatile float ratio = (float)best / den;
            win[r] = (ratio < pf) ? ignore_label : arg;
        }
        for (int64_t i = 0; i < hw; ++i) {
            int64_t v = win[reg[i]];
            if (reg[i] == 0) v = ignore_label;
            o[i] = (v == ignore_label) ? lab[i] : v;
        }
    }
    free(cnt); free(win);
    return ORACLE_OK;
}

/* ---------------------------------------------------------------------------------------
 * pseudo_selection.  pseudo_generation.py:59-93.
 *   :71     assert 0 <= mask <= 1
 *   :76-77  thr[b][c] = max_hw(mask[b][c]) * float32(cutoff_top)      (in-place float32 mul)
 *   :80-81  thr = max(thr, float32(cutoff_low))
 *   :83     passing = mask > thr   (strict)
 *   :85-88  exactly one passing class -> its index, otherwise ignore_label
 * soft is [b][c][hw] float32.
 * --------------------------------------------------------------------------------------- */
int oracle_pseudo_select(const float *soft, int64_t *out, int64_t b, int c, int64_t hw,
                         double cutoff_top, double cutoff_low, int64_t ignore_label)
{
    const float top = (float)cutoff_top, low = (float)cutoff_low;
    for (int64_t i = 0; i < b * c * hw; ++i)
        if (!(soft[i] <= 1.0f) || !(soft[i] >= 0.0f)) return ORACLE_ERR_PROB_RANGE;
    float *thr = (float *)malloc((size_t)(c > 0 ? c : 1) * sizeof(float));
    if (!thr) return ORACLE_ERR_ALLOC;
    for (int64_t img = 0; img < b; ++img) {
        const float *s = soft + img * c * hw;
        for (int k = 0; k < c; ++k) {
            float m = s[(int64_t)k * hw];
            for (int64_t i = 1; i < hw; ++i) if (s[(int64_t)k * hw + i] > m) m = s[(int64_t)k * hw + i];
            volatile float t = m * top;
            float tt = t;
            thr[k] = (tt > low) ? tt : low;
        }
        for (int64_t i = 0; i < hw; ++i) {
            int n_pass = 0, which = 0;
            for (int k = 0; k < c; ++k)
                if (s[(int64_t)k * hw + i] > thr[k]) { if (!n_pass) which = k; ++n_pass; }
            out[img * hw + i] = (n_pass == 1) ? (int64_t)which : ignore_label;
        }
    }
    free(thr);
    return ORACLE_OK;
}

/* ---------------------------------------------------------------------------------------
 * DownscaleLabel.  alignment.py:466-481.
 *   :474     ignore_label -> n_classes;  one_hot(n_classes+1)
 *   :477     avg_pool2d(kernel = scale): ratio = float32(count) / float32(scale*scale)
 *   :478     (max_ratio, argmax) over the n_classes+1 planes, first maximum wins
 *   :479-480 argmax == n_classes -> ignore;  max_ratio < float32(min_ratio) -> ignore
 * label [b][H][W] -> out [b][H/scale][W/scale]  (floor; trailing rows/cols are dropped
 * exactly as avg_pool2d without padding drops them).
 * --------------------------------------------------------------------------------------- */
int oracle_downscale_label(const int64_t *label, int64_t *out, int64_t b, int64_t H, int64_t W,
                           int scale, int n_classes, int64_t ignore_label, double min_ratio)
{
    const int64_t th = H / scale, tw = W / scale;
    const float mr = (float)min_ratio;
    int64_t *cnt = (int64_t *)malloc((size_t)(n_classes + 1) * sizeof(int64_t));
    if (!cnt) return ORACLE_ERR_ALLOC;
    for (int64_t i = 0; i < b * H * W; ++i) {
        int64_t l = label[i];
        if (l == ignore_label) l = n_classes;
        if (l < 0 || l > n_classes) { free(cnt); return ORACLE_ERR_LABEL_RANGE; }
    }
    for (int64_t img = 0; img < b; ++img)
        for (int64_t y = 0; y < th; ++y)
            for (int64_t x = 0; x < tw; ++x) {
                memset(cnt, 0, (size_t)(n_classes + 1) * sizeof(int64_t));
                for (int dy = 0; dy < scale; ++dy)
                    for (int dx = 0; dx < scale; ++dx) {
                        int64_t l = label[(img * H + y * scale + dy) * W + x * scale + dx];
                        if (l == ignore_label) l = n_classes;
                        cnt[l] += 1;
                    }
                int64_t best = -1, arg = 0;
                for (int k = 0; k <= n_classes; ++k)
                    if (cnt[k] > best) { best = cnt[k]; arg = k; }
                volatile float ratio = (float)best / (float)(scale * scale);
                int64_t v = arg;
                if (arg == n_classes) v = ignore_label;
                if (ratio < mr) v = ignore_label;
                out[(img * th + y) * tw + x] = v;
            }
    free(cnt);
    return ORACLE_OK;
}

/* ---------------------------------------------------------------------------------------
 * Per-class pixel counts of a label map (ClassBalance._local_freq / _one_hot,
 * balance.py:43-66): counts[c] for c in [0, class_num), n_valid = #(label != ignore).
 * --------------------------------------------------------------------------------------- */
int oracle_class_count(const int64_t *label, int64_t n, int class_num, int64_t ignore_label,
                       int64_t *counts, int64_t *n_valid)
{
    for (int c = 0; c < class_num; ++c) counts[c] = 0;
    *n_valid = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t l = label[i];
        if (l != ignore_label) *n_valid += 1;
        if (l == ignore_label) l = class_num;
        if (l < 0 || l > class_num) return ORACLE_ERR_LABEL_RANGE;
        if (l < class_num) counts[l] += 1;
    }
    return ORACLE_OK;
}
