/* Test-only: drives the AVOption table of libswscale_b200.so through the REFERENCE's libavutil
 * (objects of oracle/_ref/obj), the way libavfilter/vf_scale.c:273,368 does.  Prototypes are declared by
 * hand so that the file builds on a box without /root/reference. */
#include <inttypes.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "swscale_b200.h"
#include "swscale_b200_cuda.h"
#include "swscale_b200_opt.h"

int av_opt_set(void *obj, const char *name, const char *val, int search_flags);
int av_opt_set_int(void *obj, const char *name, int64_t val, int search_flags);
int av_opt_set_double(void *obj, const char *name, double val, int search_flags);
int av_opt_get(void *obj, const char *name, int search_flags, uint8_t **out_val);
int av_opt_get_int(void *obj, const char *name, int search_flags, int64_t *out_val);
const AVOption *av_opt_next(const void *obj, const AVOption *prev);
void av_opt_set_defaults(void *s);
void av_free(void *ptr);
void av_log(void *avcl, int level, const char *fmt, ...);
void av_log_set_level(int level);

static void dump(SwsContext *s)
{
    printf("ctx flags=%u dither=%d alpha=%d gamma=%d src=%dx%d/%d dst=%dx%d/%d range=%d,%d chr=%d,%d,%d,%d "
           "threads=%d intent=%d scaler=%d,%d backends=%d param=%g,%g\n",
           s->flags, s->dither, s->alpha_blend, s->gamma_flag, s->src_w, s->src_h, s->src_format,
           s->dst_w, s->dst_h, s->dst_format, s->src_range, s->dst_range,
           s->src_v_chr_pos, s->src_h_chr_pos, s->dst_v_chr_pos, s->dst_h_chr_pos,
           s->threads, s->intent, s->scaler, s->scaler_sub, s->backends, s->scaler_params[0], s->scaler_params[1]);
}

int main(void)
{
    SwsContext *s = sws_alloc_context();
    const AVOption *o = NULL;
    uint8_t *str = NULL;
    int64_t v = 0;
    if (!s)
        return 1;
    while ((o = av_opt_next(s, o)))
        printf("opt %s|%s|%d|%d|%" PRId64 "|%g|%g|%g|%d|%s\n", o->name, o->help ? o->help : "", o->offset, (int)o->type,
               o->type == AV_OPT_TYPE_DOUBLE ? 0 : o->default_val.i64, o->type == AV_OPT_TYPE_DOUBLE ? o->default_val.dbl : 0.0,
               o->min, o->max, o->flags, o->unit ? o->unit : "");
    dump(s);
    memset((char *)s + sizeof(void *) * 2, 0x5a, 64);      /* scribble over the option fields ... */
    av_opt_set_defaults(s);                                  /* ... libavutil restores the defaults */
    dump(s);
    printf("set sws_flags %d\n", av_opt_set(s, "sws_flags", "bicubic+accurate_rnd+bitexact", 0));
    printf("set sws_dither %d\n", av_opt_set(s, "sws_dither", "bayer", 0));
    printf("set src_format %d\n", av_opt_set(s, "src_format", "yuv420p10le", 0));
    printf("set dst_format %d\n", av_opt_set(s, "dst_format", "rgb48le", 0));
    printf("set srcw %d\n", av_opt_set_int(s, "srcw", 3840, 0));
    printf("set srch %d\n", av_opt_set(s, "srch", "2160", 0));
    printf("set dstw %d\n", av_opt_set_int(s, "dstw", 1920, 0));
    printf("set dsth %d\n", av_opt_set_int(s, "dsth", 1080, 0));
    printf("set param0 %d\n", av_opt_set_double(s, "param0", 0.5, 0));
    printf("set threads %d\n", av_opt_set(s, "threads", "auto", 0));
    printf("set alphablend %d\n", av_opt_set(s, "alphablend", "checkerboard", 0));
    printf("set scaler %d\n", av_opt_set(s, "scaler", "lanczos", 0));
    printf("set src_range %d\n", av_opt_set(s, "src_range", "1", 0));
    printf("set dst_v_chr_pos %d\n", av_opt_set_int(s, "dst_v_chr_pos", 128, 0));
    dump(s);
    printf("bad flag %d\n", av_opt_set(s, "sws_flags", "no_such_flag", 0) < 0);
    printf("bad range %d\n", av_opt_set_int(s, "srcw", 0, 0) < 0);
    printf("bad name %d\n", av_opt_set(s, "no_such_option", "1", 0) < 0);
    if (av_opt_get(s, "sws_flags", 0, &str) >= 0) {
        printf("get sws_flags %s\n", (const char *)str);
        av_free(str);
    }
    if (av_opt_get_int(s, "dstw", 0, &v) >= 0)
        printf("get dstw %" PRId64 "\n", v);
    /* the options reach the planner: same geometry as the reference's context for these values */
    if (sws_b200_plan_only(s) >= 0) {
        int info[32];
        sws_b200_get_info(s, info);
        printf("plan chr %dx%d -> %dx%d bpc %d,%d\n", info[8], info[9], info[10], info[11], info[12], info[13]);
    } else {
        printf("plan failed: %s\n", sws_cuda_last_error(s));
    }
    av_log_set_level(32);
    av_log(s, 32, "av_log reaches the class\n");          /* -> stderr "[swscaler @ 0x...] ..." */
    sws_free_context(&s);
    return 0;
}
