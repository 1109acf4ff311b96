"""The lines a maintainer adds to the reference tree for the CUDA hook.

Each entry is (file under libswscale/, anchor line that exists verbatim in the reference, where to
insert relative to it, inserted text).  integration/build_hooked.py applies them to scratch copies at
build time -- /root/reference stays read-only and no reference source enters this repository: only the
anchors (one line each, needed to locate the insertion point) and our own added lines live here.

Reference sites (file:line in /root/reference/libswscale):
  swscale_internal.h:701    SwsInternal grows `void *cuda_priv` next to hw_priv
  swscale_internal.h:1040   prototypes next to ff_sws_init_swscale_riscv()
  swscale.c:712-713         ff_sws_init_scale(): the per-arch init chain gains the CUDA call
  swscale.c:1163            scale_internal(): a hooked context hands the slice to ff_sws_cuda_scale()
                            instead of c->convert_unscaled() / ff_swscale()
  swscale_unscaled.c:2704   ff_get_unscaled_swscale(): per-arch chain of the special converters
  utils.c:871               sws_setColorspaceDetails() forwards the caller's tables
  utils.c:2257              sws_freeContext() drops the hook state
  graph.c:406,476,497       the legacy pass of sws_scale_frame(): one slice per frame (align = 0, as the
                            reference already does for error diffusion) and a run function for hooked contexts
"""

PATCHES = [
    ("swscale_internal.h", "    void *hw_priv; /* refstruct */", "after",
     "\n    /* B200 CUDA path (libswscale/cuda/swscale_cuda.c): non-NULL once ff_sws_init_swscale_cuda() claimed the context */\n"
     "    void *cuda_priv; /* refstruct */\n"),
    ("swscale_internal.h", "void ff_sws_init_swscale_riscv(SwsInternal *c);", "after",
     "void ff_sws_init_swscale_cuda(SwsInternal *c);\n"
     "void ff_get_unscaled_swscale_cuda(SwsInternal *c);\n"
     "void ff_sws_cuda_set_colorspace(SwsInternal *c, const int inv_table[4], int srcRange, const int table[4],\n"
     "                                int dstRange, int brightness, int contrast, int saturation);\n"
     "int ff_sws_cuda_scale(SwsInternal *c, const uint8_t *const src[], const int srcStride[], int srcSliceY,\n"
     "                      int srcSliceH, uint8_t *const dst[], const int dstStride[], int dstSliceY, int dstSliceH);\n"
     "long ff_sws_cuda_launches(const SwsInternal *c);\n"
     "long ff_sws_cuda_slices_total(void);\n"
     "const char *ff_sws_cuda_kernel(const SwsInternal *c);\n"),
    # the anchor is followed by the #endif of the per-arch chain
    ("swscale.c", "    ff_sws_init_swscale_riscv(c);", "after+1",
     "#if CONFIG_SWSCALE_CUDA\n    ff_sws_init_swscale_cuda(c);\n#endif\n"),
    ("swscale.c", "    if (c->convert_unscaled) {", "before",
     "#if CONFIG_SWSCALE_CUDA\n"
     "    if (c->cuda_priv) {\n"
     "        ret = ff_sws_cuda_scale(c, src2, srcStride2, srcSliceY_internal, srcSliceH,\n"
     "                                dst2, dstStride2, dstSliceY, dstSliceH);\n"
     "    } else\n"
     "#endif\n"),
    ("swscale_unscaled.c", "    ff_get_unscaled_swscale_aarch64(c);", "after+1",
     "#if CONFIG_SWSCALE_CUDA\n    ff_get_unscaled_swscale_cuda(c);\n#endif\n"),
    ("utils.c", "    ret = handle_formats(sws);", "before:first",
     "#if CONFIG_SWSCALE_CUDA\n"
     "    ff_sws_cuda_set_colorspace(c, inv_table, srcRange, table, dstRange, brightness, contrast, saturation);\n"
     "#endif\n"),
    ("utils.c", "    av_refstruct_unref(&c->hw_priv);", "after",
     "    av_refstruct_unref(&c->cuda_priv);\n"),
    ("graph.c", "static void run_legacy_swscale(const SwsFrame *out, const SwsFrame *in,", "before",
     "#if CONFIG_SWSCALE_CUDA\n"
     "/* a hooked context converts the whole frame with one call (the pass is registered with align = 0) */\n"
     "static void run_legacy_cuda(const SwsFrame *out, const SwsFrame *in,\n"
     "                            int y, int h, const SwsPass *pass)\n"
     "{\n"
     "    SwsContext *sws = pass->priv;\n"
     "    SwsInternal *c = sws_internal(sws);\n"
     "    av_assert1(y == 0 && h == sws->dst_h);\n"
     "    ff_sws_cuda_scale(c, (const uint8_t *const *) in->data, in->linesize, 0, sws->src_h,\n"
     "                      out->data, out->linesize, 0, sws->dst_h);\n"
     "}\n"
     "#endif\n\n"),
    ("graph.c", "        align = 0; /* disable slice threading */", "after",
     "#if CONFIG_SWSCALE_CUDA\n"
     "    if (c->cuda_priv)\n"
     "        align = 0; /* one launch per frame */\n"
     "#endif\n"),
    ("graph.c", "                                c->convert_unscaled ? run_legacy_unscaled : run_legacy_swscale,", "replace",
     "#if CONFIG_SWSCALE_CUDA\n"
     "                                c->cuda_priv ? run_legacy_cuda :\n"
     "#endif\n"
     "                                c->convert_unscaled ? run_legacy_unscaled : run_legacy_swscale,\n"),
]


def apply(name, text):
    """Return `text` of libswscale/<name> with every patch for it applied; raises if an anchor is missing."""
    for fname, anchor, where, ins in PATCHES:
        if fname != name:
            continue
        lines = text.split("\n")
        hits = [i for i, l in enumerate(lines) if l == anchor]
        if where == "before:first":
            hits = hits[:1]
            where = "before"
        if len(hits) != 1:
            raise RuntimeError("%s: anchor %r found %d times" % (name, anchor, len(hits)))
        i = hits[0]
        ins_lines = ins.rstrip("\n").split("\n")
        if where == "before":
            lines[i:i] = ins_lines
        elif where == "after":
            lines[i + 1:i + 1] = ins_lines
        elif where == "after+1":
            lines[i + 2:i + 2] = ins_lines
        elif where == "replace":
            lines[i:i + 1] = ins_lines
        else:
            raise ValueError(where)
        text = "\n".join(lines)
    return text


def patched_files():
    return sorted({p[0] for p in PATCHES})
