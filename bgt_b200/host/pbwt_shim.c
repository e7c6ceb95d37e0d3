/*
 * pbwt_shim.c -- seam A (include/pbwt_b200.h): pbf_open_r / pbf_read / pbf_seek / pbf_subset of the reference
 * (pbwt.c:221-262, 313-388) and its write side pbf_open_w / pbf_write (pbwt.c:199-219, 288-311) implemented over
 * the C ABI of libbgt_b200.so.  Host logic in C, as in the reference; every row is decoded / encoded on the GPU.
 * Reading: a window of checkpoint blocks is resident in HBM at a time and rows are decoded in batches sized by the
 * output width and the access pattern; pbf_read hands out pointers into the current batch.  Writing: rows are collected
 * into batches, encoded by b200_enc_write_bytes and their records written to the file batch by batch.
 *
 * This is library code: no entry point terminates the process.  Failures return NULL / -1 as the reference's do
 * (pbwt.c:228-235, 291, 317, 353-359) and leave a message in pbf_b200_strerror().
 */
#include <fcntl.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "../../include/bgt_b200.h"
#include "../../include/pbwt_b200.h"

#define SHIM_MAGIC 0x42323030504246ULL /* "B200PBF" */
#define SHIM_MAX_DEV 16

struct pbf_s {
	uint64_t magic;
	const uint8_t *map; size_t map_len;
	int map_owned;              /* 1: malloc'ed copy of stdin, 0: mmap */
	int32_t m, g, shift;
	int64_t n, k;               /* rows; next row to read */
	b200_pbf_t *win;            /* resident window */
	int64_t win_beg, win_end;
	b200_query_t *q;
	int n_sub; int *sub;
	int64_t bat_beg, bat_end;   /* decoded batch [bat_beg, bat_end) */
	int64_t bat_rows_next;      /* rows of the next batch: small after a seek, doubled on every sequential refill */
	uint8_t *bat[2]; size_t bat_cap;
	const uint8_t *ret[2];
	/* writer */
	int is_writing, failed;
	FILE *fp;
	b200_enc_t *enc;
	uint8_t *wrow[2]; int64_t w_n, w_cap;  /* rows collected for the next encoder batch */
};

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static b200_ctx_t *g_ctx[SHIM_MAX_DEV];
static int (*g_foreign_close)(pbf_t *);
static __thread char t_err[512];
static int64_t g_route[PBF_B200_ROUTE_SLOTS];
static int g_route_hooked;

void pbf_b200_set_foreign_close(int (*close_fn)(pbf_t *)) { g_foreign_close = close_fn; }
const char *pbf_b200_strerror(void) { return t_err; }

static void shim_err(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(t_err, sizeof(t_err), fmt, ap);
	va_end(ap);
	fprintf(stderr, "[E::bgt_b200] %s\n", t_err);
}

/* ---- route counters: which path served this process (tests assert on them; BGT_B200_ROUTE=1 prints them at exit) */
static void route_print(void)
{
	fprintf(stderr, "[b200 route] view_fast=%lld view_fast_to_ref=%lld seamB_batches=%lld ref_bgtm_read=%lld seamA_batches=%lld enc_batches=%lld gpus=%lld region_launches=%lld\n",
	        (long long)g_route[0], (long long)g_route[1], (long long)g_route[2], (long long)g_route[3], (long long)g_route[4], (long long)g_route[5],
	        (long long)g_route[6], (long long)g_route[7]);
}

/* for a process that leaves through _exit (the CLI's fast path skips the teardown): what atexit would have printed */
void pbf_b200_route_report(void)
{
	const char *e = getenv("BGT_B200_ROUTE");
	if (e && *e == '1') route_print();
}

void pbf_b200_route_add(int slot, int64_t n)
{
	if (slot < 0 || slot >= PBF_B200_ROUTE_SLOTS) return;
	if (!g_route_hooked) {
		const char *e = getenv("BGT_B200_ROUTE");
		pthread_mutex_lock(&g_lock);
		if (!g_route_hooked) { g_route_hooked = 1; if (e && *e == '1') atexit(route_print); }
		pthread_mutex_unlock(&g_lock);
	}
	if (slot == 6) g_route[slot] = n; else __sync_fetch_and_add(&g_route[slot], n);
}

int64_t pbf_b200_route_get(int slot) { return slot >= 0 && slot < PBF_B200_ROUTE_SLOTS ? g_route[slot] : -1; }

/* one context per device, created on first use; NULL (with a message) if the device cannot be used -- there is no CPU fallback */
b200_ctx_t *pbf_b200_ctx_dev(int dev)
{
	b200_ctx_t *c;
	if (dev < 0 || dev >= SHIM_MAX_DEV) { shim_err("device %d out of range", dev); return 0; }
	pthread_mutex_lock(&g_lock);
	if (g_ctx[dev] == 0) {
		g_ctx[dev] = b200_ctx_create(dev);
		if (g_ctx[dev] == 0) shim_err("%s", b200_strerror());
	}
	c = g_ctx[dev];
	pthread_mutex_unlock(&g_lock);
	return c;
}

static int default_dev(void)
{
	const char *d = getenv("BGT_B200_DEVICE");
	return d && *d ? atoi(d) : 0;
}

static b200_ctx_t *shim_ctx(void) { return pbf_b200_ctx_dev(default_dev()); }

/* used by the seam-B shim (integration/bgtm_shim.c) to share the context and the file mapping */
b200_ctx_t *pbf_b200_ctx(void) { return shim_ctx(); }
const uint8_t *pbf_b200_image(const pbf_t *pb, size_t *len)
{
	if (pb == 0 || pb->magic != SHIM_MAGIC || pb->is_writing) return 0;
	*len = pb->map_len;
	return pb->map;
}

static int64_t env_i64(const char *name, int64_t dflt)
{
	const char *s = getenv(name);
	return s && *s ? atoll(s) : dflt;
}

/* stdin (pbwt.c:227-230): the stream is read to its end; the index record at its tail is then available like a file's */
static uint8_t *slurp(FILE *fp, size_t *len)
{
	size_t cap = 1u << 20, n = 0;
	uint8_t *buf = (uint8_t*)malloc(cap);
	while (buf) {
		const size_t got = fread(buf + n, 1, cap - n, fp);
		n += got;
		if (got == 0) break;
		if (n == cap) { uint8_t *nb = (uint8_t*)realloc(buf, cap <<= 1); if (nb == 0) { free(buf); return 0; } buf = nb; }
	}
	*len = n;
	return buf;
}

pbf_t *pbf_open_r(const char *fn)
{
	pbf_t *pb;
	const uint8_t *img;
	size_t len;
	int owned = 0;
	int32_t v[3];
	uint64_t ioff;
	t_err[0] = 0;
	if (fn == 0 || strcmp(fn, "-") == 0) {
		img = slurp(stdin, &len);
		owned = 1;
		if (img == 0) return 0;
	} else {
		struct stat sb;
		void *mp;
		const int fd = open(fn, O_RDONLY);
		if (fd < 0) return 0;                                             /* pbwt.c:228-229 */
		if (fstat(fd, &sb) != 0 || sb.st_size < 16) { close(fd); return 0; }
		mp = mmap(0, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
		close(fd);
		if (mp == MAP_FAILED) return 0;
		img = (const uint8_t*)mp; len = (size_t)sb.st_size;
	}
	if (len < 16 || memcmp(img, "PBF\1", 4) != 0) goto fail;              /* pbwt.c:232-235 */
	memcpy(v, img + 4, 12);
	if (v[1] != 2) { shim_err("'%s' has %d bit planes; the B200 genotype path reads BGT's 2 (import.c:68)", fn ? fn : "-", v[1]); goto fail; }
	/* the index record (pbwt.c:247-258); the offset at the tail is untrusted: compare without adding to it */
	if (len < 16 + 13 + 8) { shim_err("'%s' has no index record", fn ? fn : "-"); goto fail; }
	memcpy(&ioff, img + len - 8, 8);
	if (ioff > len - 13 - 8 || ioff < 16 || img[ioff] != 'I') { shim_err("'%s' has no (valid) index record; it cannot be decoded block-wise", fn ? fn : "-"); goto fail; }
	pb = (pbf_t*)calloc(1, sizeof(pbf_t));
	pb->magic = SHIM_MAGIC;
	pb->map = img; pb->map_len = len; pb->map_owned = owned;
	pb->m = v[0]; pb->g = v[1]; pb->shift = v[2];
	memcpy(&pb->n, img + ioff + 1, 8);
	pb->bat_beg = pb->bat_end = -1;
	pb->win_beg = pb->win_end = -1;
	return pb;
fail:
	if (owned) free((void*)img); else munmap((void*)img, len);
	return 0;
}

/* pbwt.c:199-219.  g must be 2 (what BGT writes, import.c:68); NULL/"-" = stdout like the reference. */
pbf_t *pbf_open_w(const char *fn, int m, int g, int shift)
{
	pbf_t *pb;
	FILE *fp;
	b200_ctx_t *ctx;
	t_err[0] = 0;
	if (g != 2) { shim_err("pbf_open_w: %d bit planes; the B200 encoder writes BGT's 2 (import.c:68)", g); return 0; }
	if ((ctx = shim_ctx()) == 0) return 0;
	if (fn && strcmp(fn, "-") != 0) {
		if ((fp = fopen(fn, "wb")) == NULL) return 0;
	} else fp = stdout;
	pb = (pbf_t*)calloc(1, sizeof(pbf_t));
	pb->magic = SHIM_MAGIC; pb->is_writing = 1; pb->fp = fp;
	pb->m = m; pb->g = g; pb->shift = shift;
	pb->enc = b200_enc_create(ctx, m, shift);
	pb->w_cap = (64LL << 20) / (m > 0 ? m : 1);
	if (pb->w_cap < 16) pb->w_cap = 16;
	if (pb->w_cap > 4096) pb->w_cap = 4096;
	if (pb->enc) {
		pb->wrow[0] = (uint8_t*)b200_host_alloc((size_t)pb->w_cap * m);
		pb->wrow[1] = (uint8_t*)b200_host_alloc((size_t)pb->w_cap * m);
	}
	if (pb->enc == 0 || !pb->wrow[0] || !pb->wrow[1]) {
		shim_err("%s", b200_strerror());
		if (pb->enc) b200_enc_destroy(pb->enc);
		b200_host_free(pb->wrow[0]); b200_host_free(pb->wrow[1]);
		if (fp != stdout) fclose(fp);
		free(pb);
		return 0;
	}
	return pb;
}

/* encode the collected rows and write their records (streaming, pbwt.c:288-311 writes every row as it goes) */
static int flush_rows(pbf_t *pb)
{
	const uint8_t *bytes = 0;
	int64_t len;
	if (pb->failed) return -1;
	if (pb->w_n) {
		if (b200_enc_write_bytes(pb->enc, pb->wrow[0], pb->wrow[1], pb->w_n) != 0) { shim_err("%s", b200_strerror()); pb->failed = 1; return -1; }
		pbf_b200_route_add(5, 1);
	}
	pb->w_n = 0;
	len = b200_enc_drain(pb->enc, &bytes);
	if (len < 0 || (len > 0 && fwrite(bytes, 1, (size_t)len, pb->fp) != (size_t)len)) { shim_err("writing the PBF failed"); pb->failed = 1; return -1; }
	return 0;
}

/* pbwt.c:288-311: one row, a[g][m] bytes */
int pbf_write(pbf_t *pb, uint8_t *const*a)
{
	if (pb == 0 || pb->magic != SHIM_MAGIC || !pb->is_writing || pb->failed) return -1;   /* pbwt.c:291 */
	memcpy(pb->wrow[0] + (size_t)pb->w_n * pb->m, a[0], (size_t)pb->m);
	memcpy(pb->wrow[1] + (size_t)pb->w_n * pb->m, a[1], (size_t)pb->m);
	++pb->n;
	if (++pb->w_n == pb->w_cap) return flush_rows(pb);
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
	int ret = 0;
	if (pb == 0) return 0;
	if (pb->magic != SHIM_MAGIC) return g_foreign_close ? g_foreign_close(pb) : -1;
	if (pb->is_writing) { /* pbwt.c:268-276: the last rows, then the index record */
		const uint8_t *img = 0;
		int64_t len;
		ret = flush_rows(pb);
		if (ret == 0) {
			len = b200_enc_finish(pb->enc, &img);
			if (len < 0 || fwrite(img, 1, (size_t)len, pb->fp) != (size_t)len) { shim_err("writing the PBF index failed"); ret = -1; }
		}
		b200_enc_destroy(pb->enc);
		b200_host_free(pb->wrow[0]); b200_host_free(pb->wrow[1]);
		if (fclose(pb->fp) != 0) ret = -1;
		pb->magic = 0;
		free(pb);
		return ret;
	}
	drop_window(pb);
	b200_host_free(pb->bat[0]); b200_host_free(pb->bat[1]);
	free(pb->sub);
	if (pb->map_owned) free((void*)pb->map); else munmap((void*)pb->map, pb->map_len);
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
	if ((int64_t)k == pb->k) return 0;                                      /* pbwt.c:354 */
	if (k >= (uint64_t)pb->n) return -1;                                    /* pbwt.c:359 */
	if ((int64_t)k < pb->bat_beg || (int64_t)k >= pb->bat_end) pb->bat_rows_next = 0;   /* a jump: the next batch starts small again */
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
		int64_t rows, want, cap;
		if (ctx == 0) return 0;
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
			if (pb->win == 0) { shim_err("%s", b200_strerror()); return 0; }
			pb->win_beg = b200_pbf_row_beg(pb->win); pb->win_end = b200_pbf_row_end(pb->win);
		}
		if (pb->q == 0) {
			pb->q = b200_query_create_cols(ctx, pb->win, pb->n_sub, pb->sub);
			if (pb->q == 0) { shim_err("%s", b200_strerror()); return 0; }
		}
		/* batch size: bounded by the decoded bytes; after a seek it starts at a few rows and doubles on every sequential
		 * refill, so that scattered pbf_seek + pbf_read access (region / -a queries through the reference's bgtm_read) does
		 * not decode tens of MB of rows ahead of every landing point */
		cap = env_i64("BGT_B200_BATCH_BYTES", 64LL << 20) / (width > 0 ? width : 1);
		if (cap < 1) cap = 1;
		if (cap > 65536) cap = 65536;
		want = pb->bat_rows_next > 0 ? pb->bat_rows_next : (pb->k == 0 ? cap : 64);
		if (want > cap) want = cap;
		pb->bat_rows_next = want * 4 < cap ? want * 4 : cap;
		rows = pb->win_end - pb->k < want ? pb->win_end - pb->k : want;
		if ((size_t)rows * width > pb->bat_cap) {
			b200_host_free(pb->bat[0]); b200_host_free(pb->bat[1]);
			pb->bat_cap = (size_t)(cap < pb->n ? cap : pb->n) * width;
			if (pb->bat_cap < (size_t)rows * width) pb->bat_cap = (size_t)rows * width;
			pb->bat[0] = (uint8_t*)b200_host_alloc(pb->bat_cap);
			pb->bat[1] = (uint8_t*)b200_host_alloc(pb->bat_cap);
			if (!pb->bat[0] || !pb->bat[1]) { shim_err("%s", b200_strerror()); pb->bat_cap = 0; return 0; }
		}
		memset(&so, 0, sizeof(so));
		so.hap_bytes[0] = pb->bat[0]; so.hap_bytes[1] = pb->bat[1];
		if (b200_scan(ctx, pb->win, pb->q, pb->k, rows, B200_SCAN_HAP_BYTES, &so) != rows) { shim_err("%s", b200_strerror()); return 0; }
		pbf_b200_route_add(4, 1);
		pb->bat_beg = pb->k; pb->bat_end = pb->k + rows;
	}
	pb->ret[0] = pb->bat[0] + (size_t)(pb->k - pb->bat_beg) * width;
	pb->ret[1] = pb->bat[1] + (size_t)(pb->k - pb->bat_beg) * width;
	++pb->k;
	return pb->ret;
}

int pbf_get_g(const pbf_t *pb) { return pb->g; }
int pbf_get_m(const pbf_t *pb) { return pb->m; }
int pbf_get_n(const pbf_t *pb) { return (int)pb->n; }                      /* int in the reference's API as well (pbwt.c:392) */
int pbf_get_shift(const pbf_t *pb) { return pb->shift; }
