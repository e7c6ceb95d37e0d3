/*
 * pbwt_shim.c -- seam A (include/pbwt_b200.h): pbf_open_r / pbf_read / pbf_seek / pbf_subset of the reference
 * (pbwt.c:221-262, 313-388) and its write side pbf_open_w / pbf_write (pbwt.c:199-219, 288-311) implemented over
 * the C ABI of libbgt_b200.so.  Host logic in C, as in the reference; every row is decoded / encoded on the GPU.
 * Reading: a window of checkpoint blocks is resident in HBM at a time and rows are decoded in batches sized by the
 * output width; pbf_read hands out pointers into the current batch.  Writing: rows are collected into batches and
 * encoded by b200_enc_write_bytes; the file is written when the handle is closed.
 */
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "../../include/bgt_b200.h"
#include "../../include/pbwt_b200.h"

#define SHIM_MAGIC 0x42323030504246ULL /* "B200PBF" */

struct pbf_s {
	uint64_t magic;
	const uint8_t *map; size_t map_len;
	int32_t m, g, shift;
	int64_t n, k;               /* rows; next row to read */
	b200_pbf_t *win;            /* resident window */
	int64_t win_beg, win_end;
	b200_query_t *q;
	int n_sub; int *sub;
	int64_t bat_beg, bat_end;   /* decoded batch [bat_beg, bat_end) */
	uint8_t *bat[2]; size_t bat_cap;
	const uint8_t *ret[2];
	/* writer */
	int is_writing;
	FILE *fp;
	b200_enc_t *enc;
	uint8_t *wrow[2]; int64_t w_n, w_cap;  /* rows collected for the next encoder batch */
};

static b200_ctx_t *g_ctx;
static int (*g_foreign_close)(pbf_t *);

void pbf_b200_set_foreign_close(int (*close_fn)(pbf_t *)) { g_foreign_close = close_fn; }

static b200_ctx_t *shim_ctx(void)
{
	if (!g_ctx) {
		const char *d = getenv("BGT_B200_DEVICE");
		g_ctx = b200_ctx_create(d ? atoi(d) : 0);
		if (!g_ctx) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); exit(1); } /* no CPU fallback */
	}
	return g_ctx;
}

/* used by the seam-B shim (integration/bgtm_shim.c) to share the context and the file mapping */
b200_ctx_t *pbf_b200_ctx(void) { return shim_ctx(); }
const uint8_t *pbf_b200_image(const pbf_t *pb, size_t *len)
{
	if (pb == 0 || pb->magic != SHIM_MAGIC) return 0;
	*len = pb->map_len;
	return pb->map;
}

static int64_t env_i64(const char *name, int64_t dflt)
{
	const char *s = getenv(name);
	return s && *s ? atoll(s) : dflt;
}

pbf_t *pbf_open_r(const char *fn)
{
	int fd;
	struct stat sb;
	void *mp;
	pbf_t *pb;
	int32_t v[3];
	uint64_t ioff;
	if (fn == 0 || strcmp(fn, "-") == 0) { fprintf(stderr, "[E::bgt_b200] reading a PBF from stdin is not supported\n"); return 0; }
	if ((fd = open(fn, O_RDONLY)) < 0) return 0;                          /* pbwt.c:228-229 */
	if (fstat(fd, &sb) != 0 || sb.st_size < 16) { close(fd); return 0; }
	mp = mmap(0, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (mp == MAP_FAILED) return 0;
	if (memcmp(mp, "PBF\1", 4) != 0) { munmap(mp, (size_t)sb.st_size); return 0; } /* pbwt.c:232-235 */
	pb = (pbf_t*)calloc(1, sizeof(pbf_t));
	pb->magic = SHIM_MAGIC;
	pb->map = (const uint8_t*)mp; pb->map_len = (size_t)sb.st_size;
	memcpy(v, pb->map + 4, 12);
	pb->m = v[0]; pb->g = v[1]; pb->shift = v[2];
	if (pb->map_len >= 16 + 8) {                                          /* pbwt.c:247-258 */
		memcpy(&ioff, pb->map + pb->map_len - 8, 8);
		if (ioff + 13 <= pb->map_len && pb->map[ioff] == 'I') memcpy(&pb->n, pb->map + ioff + 1, 8);
	}
	pb->bat_beg = pb->bat_end = -1;
	pb->win_beg = pb->win_end = -1;
	return pb;
}

/* pbwt.c:199-219.  g must be 2 (what BGT writes, import.c:68); NULL/"-" = stdout like the reference. */
pbf_t *pbf_open_w(const char *fn, int m, int g, int shift)
{
	pbf_t *pb;
	FILE *fp;
	if (g != 2) { fprintf(stderr, "[E::bgt_b200] pbf_open_w: %d bit planes; the B200 encoder writes BGT's 2 (import.c:68)\n", g); return 0; }
	if (fn && strcmp(fn, "-") != 0) {
		if ((fp = fopen(fn, "wb")) == NULL) return 0;
	} else fp = stdout;
	pb = (pbf_t*)calloc(1, sizeof(pbf_t));
	pb->magic = SHIM_MAGIC; pb->is_writing = 1; pb->fp = fp;
	pb->m = m; pb->g = g; pb->shift = shift;
	pb->enc = b200_enc_create(shim_ctx(), m, shift);
	if (pb->enc == 0) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); exit(1); } /* no CPU fallback */
	pb->w_cap = (64LL << 20) / (m > 0 ? m : 1);
	if (pb->w_cap < 16) pb->w_cap = 16;
	if (pb->w_cap > 4096) pb->w_cap = 4096;
	pb->wrow[0] = (uint8_t*)b200_host_alloc((size_t)pb->w_cap * m);
	pb->wrow[1] = (uint8_t*)b200_host_alloc((size_t)pb->w_cap * m);
	if (!pb->wrow[0] || !pb->wrow[1]) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); exit(1); }
	return pb;
}

static void flush_rows(pbf_t *pb)
{
	if (pb->w_n && b200_enc_write_bytes(pb->enc, pb->wrow[0], pb->wrow[1], pb->w_n) != 0) {
		fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror());
		exit(1);
	}
	pb->w_n = 0;
}

/* pbwt.c:288-311: one row, a[g][m] bytes */
int pbf_write(pbf_t *pb, uint8_t *const*a)
{
	if (pb == 0 || pb->magic != SHIM_MAGIC || !pb->is_writing) return -1;   /* pbwt.c:291 */
	memcpy(pb->wrow[0] + (size_t)pb->w_n * pb->m, a[0], (size_t)pb->m);
	memcpy(pb->wrow[1] + (size_t)pb->w_n * pb->m, a[1], (size_t)pb->m);
	if (++pb->w_n == pb->w_cap) flush_rows(pb);
	++pb->n;
	return 0;
}

static void drop_window(pbf_t *pb)
{
	if (pb->q) { b200_query_destroy(pb->q); pb->q = 0; }
	if (pb->win) { b200_pbf_close(pb->win); pb->win = 0; }
	pb->win_beg = pb->win_end = -1;
}

int pbf_close(pbf_t *pb)
{
	if (pb == 0) return 0;
	if (pb->magic != SHIM_MAGIC) return g_foreign_close ? g_foreign_close(pb) : -1;
	if (pb->is_writing) { /* pbwt.c:268-276: the index goes out with the rest of the file */
		const uint8_t *img = 0;
		int64_t len;
		flush_rows(pb);
		len = b200_enc_finish(pb->enc, &img);
		if (len < 0 || fwrite(img, 1, (size_t)len, pb->fp) != (size_t)len) { fprintf(stderr, "[E::bgt_b200] writing the PBF failed\n"); exit(1); }
		b200_enc_destroy(pb->enc);
		b200_host_free(pb->wrow[0]); b200_host_free(pb->wrow[1]);
		fclose(pb->fp);
		pb->magic = 0;
		free(pb);
		return 0;
	}
	drop_window(pb);
	b200_host_free(pb->bat[0]); b200_host_free(pb->bat[1]);
	free(pb->sub);
	munmap((void*)pb->map, pb->map_len);
	pb->magic = 0;
	free(pb);
	return 0;
}

int pbf_subset(pbf_t *pb, int n_sub, int *sub)
{
	if (pb == 0 || pb->magic != SHIM_MAGIC) return -1;
	if (n_sub <= 0 || n_sub >= pb->m || sub == 0) n_sub = 0;                /* pbwt.c:377 */
	pb->n_sub = n_sub;
	free(pb->sub); pb->sub = 0;
	if (n_sub) {
		pb->sub = (int*)malloc((size_t)n_sub * sizeof(int));
		memcpy(pb->sub, sub, (size_t)n_sub * sizeof(int));                  /* pbwt.c:379-380 */
	}
	if (pb->q) { b200_query_destroy(pb->q); pb->q = 0; }
	pb->bat_beg = pb->bat_end = -1;
	return 0;
}

int pbf_seek(pbf_t *pb, uint64_t k)
{
	if (pb == 0 || pb->magic != SHIM_MAGIC || pb->is_writing) return -1;    /* pbwt.c:353 */
	if ((int64_t)k > pb->n) return -1;                                      /* pbwt.c:359 */
	pb->k = (int64_t)k;                                                     /* "next row to read", pbwt.c:334,354 */
	return 0;
}

const uint8_t **pbf_read(pbf_t *pb)
{
	int width;
	if (pb == 0 || pb->magic != SHIM_MAGIC || pb->is_writing) return 0;     /* pbwt.c:317 */
	if (pb->k >= pb->n) return 0;                                           /* 'I' record reached, pbwt.c:335 */
	width = pb->n_sub ? pb->n_sub : pb->m;
	if (pb->k < pb->bat_beg || pb->k >= pb->bat_end) {
		b200_ctx_t *ctx = shim_ctx();
		b200_scan_out_t so;
		const int64_t BS = 1LL << pb->shift;
		int64_t rows, want;
		if (pb->win == 0 || pb->k < pb->win_beg || pb->k >= pb->win_end) { /* make the blocks around row k resident */
			int64_t wblocks = env_i64("BGT_B200_WINDOW_BLOCKS", 0), beg = pb->k / BS * BS, end;
			if (wblocks <= 0) { /* about 1 GB of snapshots + RLE per window */
				wblocks = (1LL << 30) / (8LL * pb->m + 256LL * BS);
				if (wblocks < 1) wblocks = 1;
			}
			end = beg + wblocks * BS;
			if (end > pb->n) end = pb->n;
			drop_window(pb);
			pb->win = b200_pbf_load(ctx, pb->map, pb->map_len, beg, end);
			if (pb->win == 0) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); return 0; }
			pb->win_beg = b200_pbf_row_beg(pb->win); pb->win_end = b200_pbf_row_end(pb->win);
		}
		if (pb->q == 0) {
			pb->q = b200_query_create_cols(ctx, pb->win, pb->n_sub, pb->sub);
			if (pb->q == 0) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); return 0; }
		}
		want = env_i64("BGT_B200_BATCH_BYTES", 64LL << 20) / (width > 0 ? width : 1);
		if (want < 1) want = 1;
		if (want > 65536) want = 65536;
		rows = pb->win_end - pb->k < want ? pb->win_end - pb->k : want;
		if ((size_t)rows * width > pb->bat_cap) {
			b200_host_free(pb->bat[0]); b200_host_free(pb->bat[1]);
			pb->bat_cap = (size_t)rows * width;
			pb->bat[0] = (uint8_t*)b200_host_alloc(pb->bat_cap);
			pb->bat[1] = (uint8_t*)b200_host_alloc(pb->bat_cap);
			if (!pb->bat[0] || !pb->bat[1]) { fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror()); return 0; }
		}
		memset(&so, 0, sizeof(so));
		so.hap_bytes[0] = pb->bat[0]; so.hap_bytes[1] = pb->bat[1];
		if (b200_scan(ctx, pb->win, pb->q, pb->k, rows, B200_SCAN_HAP_BYTES, &so) != rows) {
			fprintf(stderr, "[E::bgt_b200] %s\n", b200_strerror());
			return 0;
		}
		pb->bat_beg = pb->k; pb->bat_end = pb->k + rows;
	}
	pb->ret[0] = pb->bat[0] + (size_t)(pb->k - pb->bat_beg) * width;
	pb->ret[1] = pb->bat[1] + (size_t)(pb->k - pb->bat_beg) * width;
	++pb->k;
	return pb->ret;
}

int pbf_get_g(const pbf_t *pb) { return pb->g; }
int pbf_get_m(const pbf_t *pb) { return pb->m; }
int pbf_get_n(const pbf_t *pb) { return (int)pb->n; }
int pbf_get_shift(const pbf_t *pb) { return pb->shift; }
