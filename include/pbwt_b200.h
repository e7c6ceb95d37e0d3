/*
 * pbwt_b200.h -- seam A: the reference's PBF reader API (pbwt.h:35-96) served by the B200 library.
 *
 * Same names, argument meaning and error behaviour as the reference, so that code written against pbwt.h
 * (bgt.c:98,114,242,341-342; pbfview.c:78-99) links against libpbwt_b200.so instead of pbwt.o unchanged:
 *
 *   pbf_open_r   pbwt.c:221-262   NULL on open failure / bad magic.  "-"/NULL (stdin) is not supported: the file
 *                                 is mapped and made resident in HBM block-window by block-window.
 *   pbf_subset   pbwt.c:374-388   any column list, output in list order; n_sub<=0 or >=m or sub==NULL -> all columns.
 *   pbf_seek     pbwt.c:349-372   0 on success, -1 if k is past the last row.
 *   pbf_read     pbwt.c:313-337   g pointers to one byte per (selected) column, owned by the handle and valid
 *                                 until the next call; NULL at the end of the file.
 *   pbf_get_*    pbwt.c:390-393
 *   pbf_close    pbwt.c:264-286
 *
 *   pbf_open_w   pbwt.c:199-219   g must be 2 (import.c:68); NULL/"-" = stdout.  NULL if the file cannot be created.
 *   pbf_write    pbwt.c:288-311   one row, a[g] = m bytes per plane (non-zero = 1, pbwt.c:61); rows are batched and encoded
 *                                 on the GPU (b200_enc_*); the file (header, 'S'/'B' records, index) is written at pbf_close.
 *
 * The in-memory codec entry points (pbc_*, pbs_dec) are not part of the seam; INTEGRATION.md shows how a host
 * application keeps its own objects for them.
 */
#ifndef PBWT_B200_H
#define PBWT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct pbf_s;
typedef struct pbf_s pbf_t;

pbf_t *pbf_open_r(const char *fn);
pbf_t *pbf_open_w(const char *fn, int m, int g, int shift);
int pbf_write(pbf_t *pb, uint8_t *const*a);
int pbf_close(pbf_t *pb);
const uint8_t **pbf_read(pbf_t *pb);
int pbf_seek(pbf_t *pb, uint64_t k);
int pbf_subset(pbf_t *pb, int n_sub, int *sub);
int pbf_get_g(const pbf_t *pb);
int pbf_get_m(const pbf_t *pb);
int pbf_get_n(const pbf_t *pb);
int pbf_get_shift(const pbf_t *pb);

/* optional hook: a host application that keeps its own writer registers its pbf_close
 * so that handles not created by pbf_open_r above are passed through */
void pbf_b200_set_foreign_close(int (*close_fn)(pbf_t *));

#ifdef __cplusplus
}
#endif
#endif
