/*
 * pbwt_b200.h -- seam A: the reference's PBF reader API (pbwt.h:35-96) served by the B200 library.
 *
 * Same names, argument meaning and error behaviour as the reference, so that code written against pbwt.h
 * (bgt.c:98,114,242,341-342; pbfview.c:78-99) links against libpbwt_b200.so instead of pbwt.o unchanged:
 *
 *   pbf_open_r   pbwt.c:221-262   NULL on open failure / bad magic / missing index record (with a message).  The file is
 *                                 mapped and made resident in HBM block-window by block-window; "-"/NULL = stdin, read to
 *                                 its end into memory first (pbwt.c:227-230).
 *   pbf_subset   pbwt.c:374-388   any column list, output in list order; n_sub<=0 or >=m or sub==NULL -> all columns.
 *   pbf_seek     pbwt.c:349-372   0 on success, -1 if k >= n (pbwt.c:359) or the handle is a writer.
 *   pbf_read     pbwt.c:313-337   g pointers to one byte per (selected) column, owned by the handle and valid
 *                                 until the next call; NULL at the end of the file.
 *   pbf_get_*    pbwt.c:390-393
 *   pbf_close    pbwt.c:264-286
 *
 *   pbf_open_w   pbwt.c:199-219   g must be 2 (import.c:68); NULL/"-" = stdout.  NULL if the file cannot be created.
 *   pbf_write    pbwt.c:288-311   one row, a[g] = m bytes per plane (non-zero = 1, pbwt.c:61); rows are batched, encoded on the
 *                                 GPU (b200_enc_*) and their records written to the file batch by batch (streaming, like the
 *                                 reference); pbf_close writes the last batch and the index.  -1 on an encoder / write error.
 *
 *   pbc_init, pbc_enc, pbc_dec, pbs_dec, pbc_enc_core, pbc_dec_core   pbwt.h:98-130, pbwt.c:57-170: the one-row in-memory codec on
 *                                 the caller's pbc_t, host C (bgt_b200/host/pbc_host.c; no caller on the hot path) -- with them
 *                                 libpbwt_b200.so replaces pbwt.o entirely.
 *
 * No entry point terminates the process: failures return NULL / -1 like the reference's and leave a message in
 * pbf_b200_strerror() (also printed to stderr once).
 */
#ifndef PBWT_B200_H
#define PBWT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PBWT_H /* the reference's pbwt.h, if the host application includes it as well, declares the same types */
typedef struct { /* full codec, pbwt.h:6-9 */
	int32_t m, l, *S0, *S;
	uint8_t *u;
} pbc_t;

typedef struct { /* pbwt.h:11-14 */
	uint32_t r;
	uint32_t i;
} pbs_dat_t;

struct pbf_s;
typedef struct pbf_s pbf_t;
#endif

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

pbc_t *pbc_init(int m);
void pbc_enc(pbc_t *pb, const uint8_t *a);
void pbc_dec(pbc_t *pb, const uint8_t *b);
void pbs_dec(int m, int n_sub, pbs_dat_t *sub, const uint8_t *u, uint8_t *a);
int  pbc_enc_core(int m, const int32_t *S0, const uint8_t *a, int32_t *S, uint8_t *u);
void pbc_dec_core(int m, const int32_t *S0, const uint8_t *u, int32_t *S, uint8_t *a);

/* last error message of the seam (empty if none) */
const char *pbf_b200_strerror(void);
/* Which path served the calls of this process (tests and `BGT_B200_ROUTE=1`, which prints the counters to stderr at exit):
 * slots: 0 device `view` pipelines run (view_fast.c), 1 of those that handed over to the reference's main_view, 2 seam-B scan
 * batches (bgtm_shim.c), 3 records served by the reference's own bgtm_read, 4 seam-A decode batches (pbf_read), 5 seam-A
 * encoder batches (pbf_write), 6 GPUs used by the last device `view`, 7 seam-B batched-region launches. */
#define PBF_B200_ROUTE_SLOTS 8
void    pbf_b200_route_add(int slot, int64_t n);
int64_t pbf_b200_route_get(int slot);
void    pbf_b200_route_report(void);   /* print them now if BGT_B200_ROUTE=1 (for processes that leave through _exit) */

/* optional hook: a host application that keeps its own writer registers its pbf_close
 * so that handles not created by pbf_open_r above are passed through */
void pbf_b200_set_foreign_close(int (*close_fn)(pbf_t *));

#ifdef __cplusplus
}
#endif
#endif
