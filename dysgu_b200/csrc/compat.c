/*
 * libdysgu_b200_compat.so - the reference's own symbol names over libdysgu_b200_align.so.
 *
 * dysgu's two Cython wrappers bind the C cores by name: _ssw_wrapper.pyx:19-48 declares ssw_init / ssw_align /
 * init_destroy / align_destroy from "ssw.h" (ssw.h:72-125) and edlib.pxd:7-38 declares edlibAlign /
 * edlibNewAlignConfig / edlibDefaultAlignConfig / edlibFreeAlignResult / edlibAlignmentToCigar from "edlib.h"
 * (edlib.h:140-271); meson.build:177-193 links ssw.c / edlib.cpp into the two extension modules.  Linking the same
 * unmodified .pyx files against this library instead (-ldysgu_b200_compat, no ssw.c / edlib.cpp) gives the reference's
 * wrappers with the GPU engine behind them.  The aliases live in a library of their own so that the unprefixed names
 * never meet the reference cores (oracle/_ref) inside one link namespace.
 *
 * The structs are layout-identical (include/dysgu_b200_align.h), so every function is a plain forward.
 */
#include "../../include/dysgu_b200_align.h"

#define DB200_EXPORT __attribute__((visibility("default")))

/* ---- ssw.h:72-125 ---------------------------------------------------------------------------------------- */
DB200_EXPORT void *ssw_init(const int8_t *read, const int32_t readLen, const int8_t *mat, const int32_t n, const int8_t score_size)
{
    return db200_ssw_init(read, readLen, mat, n, score_size);
}

DB200_EXPORT void init_destroy(void *p) { db200_init_destroy((db200_s_profile *)p); }

DB200_EXPORT void *ssw_align(const void *prof, const int8_t *ref, int32_t refLen, const uint8_t weight_gapO, const uint8_t weight_gapE,
                             const uint8_t flag, const uint16_t filters, const int32_t filterd, const int32_t maskLen)
{
    return db200_ssw_align((const db200_s_profile *)prof, ref, refLen, weight_gapO, weight_gapE, flag, filters, filterd, maskLen);
}

DB200_EXPORT void align_destroy(void *a) { db200_align_destroy((db200_s_align *)a); }

/* ---- edlib.h:140-271 --------------------------------------------------------------------------------------- */
DB200_EXPORT db200_EdlibAlignConfig edlibNewAlignConfig(int k, int mode, int task, const db200_EdlibEqualityPair *additionalEqualities,
                                                        int additionalEqualitiesLength)
{
    return db200_edlibNewAlignConfig(k, mode, task, additionalEqualities, additionalEqualitiesLength);
}

DB200_EXPORT db200_EdlibAlignConfig edlibDefaultAlignConfig(void) { return db200_edlibDefaultAlignConfig(); }

DB200_EXPORT db200_EdlibAlignResult edlibAlign(const char *query, int queryLength, const char *target, int targetLength,
                                               const db200_EdlibAlignConfig config)
{
    return db200_edlibAlign(query, queryLength, target, targetLength, config);
}

DB200_EXPORT void edlibFreeAlignResult(db200_EdlibAlignResult result) { db200_edlibFreeAlignResult(result); }

DB200_EXPORT char *edlibAlignmentToCigar(const unsigned char *alignment, int alignmentLength, int cigarFormat)
{
    return db200_edlibAlignmentToCigar(alignment, alignmentLength, cigarFormat);
}
