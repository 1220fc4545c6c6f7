/* TEST INFRASTRUCTURE ONLY - see ed_oracle.c. */
#ifndef DYSGU_B200_ED_ORACLE_H
#define DYSGU_B200_ED_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* numeric values follow the reference enums (edlib.h: EdlibAlignMode / EdlibAlignTask) */
#define ED_MODE_NW 0
#define ED_MODE_SHW 1
#define ED_MODE_HW 2
#define ED_TASK_DISTANCE 0
#define ED_TASK_LOC 1
#define ED_TASK_PATH 2

typedef struct {
    int32_t status;
    int32_t editDistance;
    int32_t *endLocations;
    int32_t *startLocations;
    int32_t numLocations;
    int32_t alphabetLength;
} ed_oracle_result;

int ed_oracle_align(const char *query, int32_t m, const char *target, int32_t n,
                    int32_t mode, int32_t task, int32_t k,
                    const char *eq_pairs, int32_t n_eq_pairs, ed_oracle_result *r);
void ed_oracle_free(ed_oracle_result *r);
int ed_oracle_batch(const char *qbuf, const int64_t *qoff, const char *tbuf, const int64_t *toff,
                    int64_t npairs, int32_t mode, int32_t task, const int32_t *k,
                    int32_t *out_fixed, int32_t *loc_buf, int64_t loc_cap, int64_t *loc_off);
/* task "path" (ed_path_oracle.c): the fields above plus the edit operations of the first location's alignment */
int ed_path_oracle_align(const char *query, int32_t m, const char *target, int32_t n, int32_t mode, int32_t k,
                         const char *eq_pairs, int32_t n_eq_pairs, ed_oracle_result *r, unsigned char *ops, int32_t *n_ops);
int ed_path_oracle_batch(const char *qbuf, const int64_t *qoff, const char *tbuf, const int64_t *toff, int64_t npairs, int32_t mode,
                         const int32_t *k, const char *eq_pairs, int32_t n_eq_pairs, int32_t *out_fixed, int32_t *loc_buf,
                         int64_t loc_cap, int64_t *loc_off, unsigned char *path_buf, int32_t *path_len);
#ifdef __cplusplus
}
#endif
#endif
