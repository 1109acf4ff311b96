/*
 * sws_filter.c -- host-side construction of the FIR banks the kernels consume.
 *
 * Restates the arithmetic of the reference's initFilter()
 * (libswscale/utils.c:197-612) so that coefficients and start positions are
 * bit-identical: int64 fixed point for point/bilinear/bicubic/area, libm
 * double for lanczos/gauss/sinc/spline, near-zero trimming, border folding and
 * error-diffused normalisation.  Structured as a small pipeline
 * (generate -> trim -> fold borders -> normalise) rather than one function.
 *
 * Deliberate differences (all result-neutral under SWS_BITEXACT):
 *  - filterAlign is 1 (the "generic arch" choice, utils.c:1675-1679,1708-1710);
 *    x86 pads to 4/2 with zero taps, which cannot change any output;
 *  - none else: SwsFilter pre/post vectors are convolved into the rows like utils.c:385-413 (stage 1b below).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>

#include "sws_internal.h"

#define MAX_TAPS_BEFORE_CASCADE 256   /* MAX_FILTER_SIZE*16/16, utils.c:492-496 (generic APCK_SIZE) */

typedef struct Taps {
    int64_t *w;     /* [n][width] */
    int32_t *pos;   /* [n] */
    int n, width;
} Taps;

static int ilog2(unsigned v)
{
    int r = 0;
    while (v >>= 1)
        r++;
    return r;
}

static int64_t iabs64(int64_t v) { return v < 0 ? -v : v; }

/* natural cubic spline helper, utils.c:155-166 */
static double spline_piece(double a, double b, double c, double d, double dist)
{
    while (dist > 1.0) {
        double nb = b + 2.0 * c + 3.0 * d;
        double nc = c + 3.0 * d;
        double nd = -b - 3.0 * c - 6.0 * d;
        a = 0.0; b = nb; c = nc; d = nd;
        dist -= 1.0;
    }
    return ((d * dist + c) * dist + b) * dist + a;
}

static int size_factor_of(int scaler, const double *param)
{
    switch (scaler) {
    case SWS_AREA:     return 1;
    case SWS_BICUBIC:  return 4;
    case SWS_BILINEAR: return 2;
    case SWS_GAUSS:    return 8;
    case SWS_SINC:     return 20;
    case SWS_SPLINE:   return 20;
    case SWS_X:        return 8;
    case SWS_LANCZOS:  return param[0] != SWS_PARAM_DEFAULT ? (int)ceil(2 * param[0]) : 6;
    }
    return -1;
}

/* One weight of the general branch, utils.c:303-380.  d is the 2^30-scaled distance. */
static int64_t kernel_weight(const SwsFirSpec *s, int64_t d, int64_t fone)
{
    const double fd = d * (1.0 / (1 << 30));
    const double *param = s->param;
    int64_t coeff;

    switch (s->scaler) {
    case SWS_BICUBIC: {
        int64_t B = (param[0] != SWS_PARAM_DEFAULT ? param[0] :   0) * (1 << 24);
        int64_t C = (param[1] != SWS_PARAM_DEFAULT ? param[1] : 0.6) * (1 << 24);
        if (d >= 1LL << 31) {
            coeff = 0;
        } else {
            int64_t dd  = (d * d) >> 30;
            int64_t ddd = (dd * d) >> 30;
            if (d < 1LL << 30)
                coeff = (12 * (1 << 24) - 9 * B - 6 * C) * ddd +
                        (-18 * (1 << 24) + 12 * B + 6 * C) * dd +
                        (6 * (1 << 24) - 2 * B) * (1 << 30);
            else
                coeff = (-B - 6 * C) * ddd + (6 * B + 30 * C) * dd +
                        (-12 * B - 48 * C) * d + (8 * B + 24 * C) * (1 << 30);
        }
        return coeff / ((1LL << 54) / fone);
    }
    case SWS_X: {
        double A = param[0] != SWS_PARAM_DEFAULT ? param[0] : 1.0;
        double c = fd < 1.0 ? cos(fd * M_PI) : -1.0;
        c = c < 0.0 ? -pow(-c, A) : pow(c, A);
        return (int64_t)((c * 0.5 + 0.5) * fone);
    }
    case SWS_AREA: {
        int64_t d2 = d - (1 << 29);
        if (d2 * s->inc < -(1LL << (29 + 16)))
            coeff = 1.0 * (1LL << (30 + 16));
        else if (d2 * s->inc < (1LL << (29 + 16)))
            coeff = -d2 * s->inc + (1LL << (29 + 16));
        else
            coeff = 0;
        return coeff * (fone >> (30 + 16));
    }
    case SWS_GAUSS: {
        double p = param[0] != SWS_PARAM_DEFAULT ? param[0] : 3.0;
        return (int64_t)(exp2(-p * fd * fd) * fone);
    }
    case SWS_SINC:
        return (int64_t)((d ? sin(fd * M_PI) / (fd * M_PI) : 1.0) * fone);
    case SWS_LANCZOS: {
        double p = param[0] != SWS_PARAM_DEFAULT ? param[0] : 3.0;
        coeff = (int64_t)((d ? sin(fd * M_PI) * sin(fd * M_PI / p) /
                               (fd * fd * M_PI * M_PI / p) : 1.0) * fone);
        return fd > p ? 0 : coeff;
    }
    case SWS_BILINEAR:
        coeff = (1 << 30) - d;
        if (coeff < 0)
            coeff = 0;
        return coeff * (fone >> 30);
    case SWS_SPLINE: {
        double p = -2.196152422706632;
        return (int64_t)(spline_piece(1.0, 0.0, p, -p - 1.0, fd) * fone);
    }
    }
    return 0;
}

/* Stage 1: raw int64 weights + first-tap positions (utils.c:219-383). */
static int generate(Taps *t, const SwsFirSpec *s, int64_t fone)
{
    const int n = s->dst_len;
    const int64_t inc = s->inc;
    int width, i, j;

    t->n = n;
    t->pos = malloc(sizeof(*t->pos) * (size_t)(n + 3));
    if (!t->pos)
        return AVERROR(ENOMEM);

    if (iabs64(inc - 0x10000) < 10 && s->src_pos == s->dst_pos) {
        /* same sampling grid: identity */
        width = 1;
        t->w = calloc((size_t)n, sizeof(*t->w));
        if (!t->w)
            return AVERROR(ENOMEM);
        for (i = 0; i < n; i++) {
            t->w[i]   = fone;
            t->pos[i] = i;
        }
    } else if (s->scaler == SWS_POINT) {
        int64_t x = ((s->dst_pos * inc) >> 8) - ((s->src_pos * 0x8000LL) >> 7);
        width = 1;
        t->w = malloc(sizeof(*t->w) * (size_t)n);
        if (!t->w)
            return AVERROR(ENOMEM);
        for (i = 0; i < n; i++) {
            t->pos[i] = (int)((x + (1 << 15)) >> 16);
            t->w[i]   = fone;
            x += inc;
        }
    } else if ((inc <= (1 << 16) && s->scaler == SWS_AREA) || s->scaler == SWS_FAST_BILINEAR) {
        /* area upscale degenerates to 2-tap linear interpolation; SWS_FAST_BILINEAR takes the same
         * branch for every ratio (utils.c:244-245) */
        int64_t x = ((s->dst_pos * inc) >> 8) - ((s->src_pos * 0x8000LL) >> 7);
        width = 2;
        t->w = malloc(sizeof(*t->w) * (size_t)n * 2);
        if (!t->w)
            return AVERROR(ENOMEM);
        for (i = 0; i < n; i++) {
            int xx = (int)((x - (1 << 15) + (1 << 15)) >> 16);
            t->pos[i] = xx;
            for (j = 0; j < 2; j++) {
                int64_t c = fone - iabs64((int64_t)xx * (1 << 16) - x) * (fone >> 16);
                t->w[i * 2 + j] = c < 0 ? 0 : c;
                xx++;
            }
            x += inc;
        }
    } else {
        int sf = size_factor_of(s->scaler, s->param);
        int64_t x;
        if (sf <= 0 || sf > 50)
            return AVERROR(EINVAL);
        if (inc <= 1 << 16)
            width = 1 + sf;
        else
            width = 1 + (int)(((int64_t)sf * s->src_len + s->dst_len - 1) / s->dst_len);
        if (width > s->src_len - 2)
            width = s->src_len - 2;
        if (width < 1)
            width = 1;
        t->w = malloc(sizeof(*t->w) * (size_t)n * width);
        if (!t->w)
            return AVERROR(ENOMEM);
        x = ((s->dst_pos * inc) >> 7) - ((s->src_pos * 0x10000LL) >> 7);
        for (i = 0; i < n; i++) {
            int xx = (int)((x - (width - 2) * (1LL << 16)) / (1 << 17));
            t->pos[i] = xx;
            for (j = 0; j < width; j++) {
                int64_t d = iabs64(((int64_t)xx * (1 << 17)) - x) << 13;
                if (inc > 1 << 16)
                    d = d * s->dst_len / s->src_len;
                t->w[(size_t)i * width + j] = kernel_weight(s, d, fone);
                xx++;
            }
            x += 2 * inc;
        }
    }
    t->width = width;
    return 0;
}

/* Stage 1b: convolve every row with the source-side SwsFilter vector and widen it for the destination-side one
 * (utils.c:385-413).  The reference accumulates `double coeff * int64 weight` into int64 cells, i.e. it
 * truncates toward zero after every single addition; the same is done here. */
static int apply_vectors(Taps *t, const SwsFirSpec *s)
{
    const int width = t->width;
    int width2 = width;
    if (s->src_vec && s->src_vec_len > 0)
        width2 += s->src_vec_len - 1;
    if (s->dst_vec_len > 0)
        width2 += s->dst_vec_len - 1;
    if (width2 == width && !(s->src_vec && s->src_vec_len > 0))
        return 0;
    int64_t *w2 = calloc((size_t)t->n * width2, sizeof(*w2));
    if (!w2)
        return AVERROR(ENOMEM);
    for (int i = 0; i < t->n; i++) {
        const int64_t *row = t->w + (size_t)i * width;
        int64_t *out = w2 + (size_t)i * width2;
        if (s->src_vec && s->src_vec_len > 0) {
            for (int k = 0; k < s->src_vec_len; k++)
                for (int j = 0; j < width; j++)
                    out[k + j] = (int64_t)((double)out[k + j] + s->src_vec[k] * (double)row[j]);
        } else {
            for (int j = 0; j < width; j++)
                out[j] = row[j];
        }
        t->pos[i] += (width - 1) / 2 - (width2 - 1) / 2;
    }
    free(t->w);
    t->w = w2;
    t->width = width2;
    return 0;
}

/* Stage 2: shift away near-zero leading taps, measure the widest useful row
 * (utils.c:417-457).  Returns the minimal tap count. */
static int trim(Taps *t, int64_t fone)
{
    const int width = t->width;
    const double limit = SWS_MAX_REDUCE_CUTOFF * fone;
    int min_width = 0;

    for (int i = t->n - 1; i >= 0; i--) {
        int64_t *row = t->w + (size_t)i * width;
        int64_t acc = 0;
        int keep = width;

        for (int j = 0; j < width; j++) {
            acc += iabs64(row[0]);
            if (acc > limit)
                break;
            /* positions must stay monotonic */
            if (i < t->n - 1 && t->pos[i] >= t->pos[i + 1])
                break;
            memmove(row, row + 1, sizeof(*row) * (size_t)(width - 1));
            row[width - 1] = 0;
            t->pos[i]++;
        }

        acc = 0;
        for (int j = width - 1; j > 0; j--) {
            acc += iabs64(row[j]);
            if (acc > limit)
                break;
            keep--;
        }
        if (keep > min_width)
            min_width = keep;
    }
    return min_width;
}

/* Stage 3: copy into the final width and fold taps that fall outside the
 * source onto the border sample (utils.c:504-560). */
static int64_t *narrow_and_fold(Taps *t, int new_width, int min_width, int src_len, unsigned flags)
{
    const int old = t->width;
    int64_t *f = malloc(sizeof(*f) * (size_t)t->n * new_width);
    if (!f)
        return NULL;

    for (int i = 0; i < t->n; i++) {
        int64_t *row = f + (size_t)i * new_width;
        for (int j = 0; j < new_width; j++) {
            row[j] = j < old ? t->w[(size_t)i * old + j] : 0;
            if ((flags & SWS_BITEXACT) && j >= min_width)
                row[j] = 0;
        }

        if (t->pos[i] < 0) {
            for (int j = 1; j < new_width; j++) {
                int left = j + t->pos[i];
                if (left < 0)
                    left = 0;
                row[left] += row[j];
                row[j]     = 0;
            }
            t->pos[i] = 0;
        }

        if (t->pos[i] + new_width > src_len) {
            int over  = new_width - src_len;
            int shift = t->pos[i] + (over < 0 ? over : 0);
            int64_t acc = 0;
            for (int j = new_width - 1; j >= 0; j--) {
                if (t->pos[i] + j >= src_len) {
                    acc   += row[j];
                    row[j] = 0;
                }
            }
            for (int j = new_width - 1; j >= 0; j--)
                row[j] = j < shift ? 0 : row[j - shift];
            t->pos[i] -= shift;
            row[src_len - 1 - t->pos[i]] += acc;
        }
        if (t->pos[i] < 0 || t->pos[i] >= src_len) {
            free(f);
            return NULL;
        }
    }
    return f;
}

/* Stage 4: scale every row to sum to `one`, diffusing the rounding error
 * left to right (utils.c:569-588). */
static void normalise(int16_t *dst, const int64_t *f, int n, int width, int one)
{
    for (int i = 0; i < n; i++) {
        const int64_t *row = f + (size_t)i * width;
        int64_t sum = 0, err = 0;
        for (int j = 0; j < width; j++)
            sum += row[j];
        sum = (sum + one / 2) / one;
        if (!sum)
            sum = 1;
        for (int j = 0; j < width; j++) {
            int64_t v = row[j] + err;
            int64_t half = sum >> 1;
            int q = (int)((v >= 0 ? v + half : v - half) / sum);
            dst[(size_t)i * width + j] = (int16_t)q;
            err = v - q * sum;
        }
    }
}

int ff_b200_build_fir(SwsFirBank *out, const SwsFirSpec *s)
{
    Taps t = { 0 };
    int64_t *folded = NULL;
    int ratio_log = ilog2((unsigned)(s->src_len / s->dst_len > 0 ? s->src_len / s->dst_len : 1));
    const int64_t fone = 1LL << (54 - (ratio_log < 8 ? ratio_log : 8));
    int ret, min_width, width;

    memset(out, 0, sizeof(*out));
    if (s->src_len / s->dst_len == 0)
        ratio_log = 0;

    ret = generate(&t, s, fone);
    if (ret < 0)
        goto done;

    ret = apply_vectors(&t, s);
    if (ret < 0)
        goto done;

    min_width = trim(&t, fone);
    if (min_width <= 0) {
        ret = AVERROR(EINVAL);
        goto done;
    }
    width = min_width; /* filterAlign == 1 */
    if (width >= MAX_TAPS_BEFORE_CASCADE) {
        ret = SWS_B200_USE_CASCADE;
        goto done;
    }

    folded = narrow_and_fold(&t, width, min_width, s->src_len, s->flags);
    if (!folded) {
        ret = AVERROR(EINVAL);
        goto done;
    }

    out->coef = calloc((size_t)(s->dst_len + 3) * width, sizeof(*out->coef));
    out->pos  = malloc(sizeof(*out->pos) * (size_t)(s->dst_len + 3));
    if (!out->coef || !out->pos) {
        ret = AVERROR(ENOMEM);
        goto done;
    }
    normalise(out->coef, folded, s->dst_len, width, s->one);
    memcpy(out->pos, t.pos, sizeof(*out->pos) * (size_t)s->dst_len);
    out->size = width;
    out->len  = s->dst_len;
    ret = 0;

done:
    free(t.w);
    free(t.pos);
    free(folded);
    if (ret < 0)
        ff_b200_free_fir(out);
    return ret;
}

void ff_b200_free_fir(SwsFirBank *b)
{
    free(b->coef);
    free(b->pos);
    memset(b, 0, sizeof(*b));
}
