/* TEST INFRASTRUCTURE ONLY - see sw_oracle.c. */
#ifndef DYSGU_B200_SW_ORACLE_H
#define DYSGU_B200_SW_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SW_ORACLE_ERR_ARG -1
#define SW_ORACLE_ERR_OVERFLOW -2
#define SW_ORACLE_ERR_TRACEBACK -3

typedef struct {
    uint16_t score1, score2;
    int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
    uint32_t *cigar;
    int32_t cigar_len;
    int32_t used_word;
} sw_oracle_result;

int sw_oracle_align(const int8_t *read, int32_t readLen, const int8_t *ref, int32_t refLen,
                    const int8_t *mat, int32_t n, int32_t score_size, int32_t gapO, int32_t gapE,
                    int32_t flag, int32_t filters, int32_t filterd, int32_t maskLen, sw_oracle_result *r);
void sw_oracle_free(sw_oracle_result *r);
void sw_oracle_encode_nt(const char *s, int32_t len, int8_t *out);
void sw_oracle_nt_matrix(int match, int mismatch, int8_t *mat25);
int sw_oracle_batch_nt(const char *qbuf, const int64_t *qoff, const char *tbuf, const int64_t *toff,
                       int64_t npairs, int match, int mismatch, int gapO, int gapE, int score_size,
                       int flag, int filters, int filterd, const int32_t *maskLen,
                       int32_t *out_fixed, uint32_t *cigar_buf, int64_t cigar_cap, int64_t *cigar_off,
                       int32_t *status);
#ifdef __cplusplus
}
#endif
#endif
