/*
 * swscale_b200_prefix.h -- symbol prefixing for IN-TREE builds.
 *
 * When the B200 library is linked into the same image as the reference's own libswscale (the
 * ff_sws_init_swscale_cuda() hook, integration/swscale_cuda.c), the 40 entry points both libraries
 * define (reference libswscale/libswscale.v + swscale.h) would clash.  Compile the B200 host sources with
 *     -DSWS_B200_PREFIX=b200_
 * and every one of them is renamed b200_<name>; the sws_cuda_*, sws_b200_* extension entries keep their
 * names (the reference has none of them).  The stand-alone drop-in of section A does not define the macro.
 */
#ifndef SWSCALE_B200_PREFIX_H
#define SWSCALE_B200_PREFIX_H
#ifdef SWS_B200_PREFIX
#define SWS_B200_CAT2(a, b) a##b
#define SWS_B200_CAT(a, b)  SWS_B200_CAT2(a, b)
#define SWS_B200_NAME(n)    SWS_B200_CAT(SWS_B200_PREFIX, n)
#define swscale_version                        SWS_B200_NAME(swscale_version)
#define swscale_configuration                  SWS_B200_NAME(swscale_configuration)
#define swscale_license                        SWS_B200_NAME(swscale_license)
#define sws_get_class                          SWS_B200_NAME(sws_get_class)
#define sws_alloc_context                      SWS_B200_NAME(sws_alloc_context)
#define sws_free_context                       SWS_B200_NAME(sws_free_context)
#define sws_init_context                       SWS_B200_NAME(sws_init_context)
#define sws_freeContext                        SWS_B200_NAME(sws_freeContext)
#define sws_getContext                         SWS_B200_NAME(sws_getContext)
#define sws_getCachedContext                   SWS_B200_NAME(sws_getCachedContext)
#define sws_scale                              SWS_B200_NAME(sws_scale)
#define sws_getCoefficients                    SWS_B200_NAME(sws_getCoefficients)
#define sws_setColorspaceDetails               SWS_B200_NAME(sws_setColorspaceDetails)
#define sws_getColorspaceDetails               SWS_B200_NAME(sws_getColorspaceDetails)
#define sws_isSupportedInput                   SWS_B200_NAME(sws_isSupportedInput)
#define sws_isSupportedOutput                  SWS_B200_NAME(sws_isSupportedOutput)
#define sws_isSupportedEndiannessConversion    SWS_B200_NAME(sws_isSupportedEndiannessConversion)
#define sws_test_format                        SWS_B200_NAME(sws_test_format)
#define sws_test_hw_format                     SWS_B200_NAME(sws_test_hw_format)
#define sws_test_colorspace                    SWS_B200_NAME(sws_test_colorspace)
#define sws_test_primaries                     SWS_B200_NAME(sws_test_primaries)
#define sws_test_transfer                      SWS_B200_NAME(sws_test_transfer)
#define sws_test_frame                         SWS_B200_NAME(sws_test_frame)
#define sws_allocVec                           SWS_B200_NAME(sws_allocVec)
#define sws_getGaussianVec                     SWS_B200_NAME(sws_getGaussianVec)
#define sws_scaleVec                           SWS_B200_NAME(sws_scaleVec)
#define sws_normalizeVec                       SWS_B200_NAME(sws_normalizeVec)
#define sws_freeVec                            SWS_B200_NAME(sws_freeVec)
#define sws_getDefaultFilter                   SWS_B200_NAME(sws_getDefaultFilter)
#define sws_freeFilter                         SWS_B200_NAME(sws_freeFilter)
#define sws_convertPalette8ToPacked32          SWS_B200_NAME(sws_convertPalette8ToPacked32)
#define sws_convertPalette8ToPacked24          SWS_B200_NAME(sws_convertPalette8ToPacked24)
#define sws_frame_setup                        SWS_B200_NAME(sws_frame_setup)
#define sws_is_noop                            SWS_B200_NAME(sws_is_noop)
#define sws_scale_frame                        SWS_B200_NAME(sws_scale_frame)
#define sws_frame_start                        SWS_B200_NAME(sws_frame_start)
#define sws_frame_end                          SWS_B200_NAME(sws_frame_end)
#define sws_send_slice                         SWS_B200_NAME(sws_send_slice)
#define sws_receive_slice                      SWS_B200_NAME(sws_receive_slice)
#define sws_receive_slice_alignment            SWS_B200_NAME(sws_receive_slice_alignment)
#endif /* SWS_B200_PREFIX */
#endif
