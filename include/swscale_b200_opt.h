/*
 * swscale_b200_opt.h -- ABI mirrors of libavutil's AVClass / AVOption (reference libavutil/log.h:76-176,
 * opt.h:250-330,428-479), so that SwsContext.av_class is a REAL class: libavutil's av_opt_set(),
 * av_opt_set_defaults(), av_opt_get(), av_opt_next() and av_log() work on a context of this library
 * exactly as on the reference's (libavfilter/vf_scale.c:273,368 sets "sws_flags", "threads", ... by name).
 * libavutil is not part of this repository; an in-tree build includes the real headers first and these
 * declarations vanish.  tests/test_options_cpu.py links the reference's own libavutil against this
 * library and drives the option table through it.
 */
#ifndef SWSCALE_B200_OPT_H
#define SWSCALE_B200_OPT_H

#include <stdint.h>

#ifndef AVUTIL_OPT_H
enum AVOptionType {
    AV_OPT_TYPE_FLAGS = 1, AV_OPT_TYPE_INT, AV_OPT_TYPE_INT64, AV_OPT_TYPE_DOUBLE, AV_OPT_TYPE_FLOAT,
    AV_OPT_TYPE_STRING, AV_OPT_TYPE_RATIONAL, AV_OPT_TYPE_BINARY, AV_OPT_TYPE_DICT, AV_OPT_TYPE_UINT64,
    AV_OPT_TYPE_CONST, AV_OPT_TYPE_IMAGE_SIZE, AV_OPT_TYPE_PIXEL_FMT, AV_OPT_TYPE_SAMPLE_FMT,
    AV_OPT_TYPE_VIDEO_RATE, AV_OPT_TYPE_DURATION, AV_OPT_TYPE_COLOR, AV_OPT_TYPE_BOOL,
};
#define AV_OPT_FLAG_ENCODING_PARAM (1 << 0)
#define AV_OPT_FLAG_VIDEO_PARAM    (1 << 4)

typedef struct AVOption {
    const char *name;
    const char *help;
    int offset;
    enum AVOptionType type;
    union {
        int64_t i64;
        double dbl;
        const char *str;
        struct { int num, den; } q;
        const void *arr;
    } default_val;
    double min, max;
    int flags;
    const char *unit;
} AVOption;
#endif

#ifndef AVUTIL_LOG_H
#define AV_CLASS_CATEGORY_SWSCALER 9
typedef struct AVClass {
    const char *class_name;
    const char *(*item_name)(void *ctx);
    const struct AVOption *option;
    int version;
    int log_level_offset_offset;
    int parent_log_context_offset;
    int category;                                   /* AVClassCategory */
    int (*get_category)(void *ctx);
    int (*query_ranges)(void **ranges, void *obj, const char *key, int flags);
    void *(*child_next)(void *obj, void *prev);
    const struct AVClass *(*child_class_iterate)(void **iter);
    int state_flags_offset;
} AVClass;
#endif

#endif
