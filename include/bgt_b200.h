/*
 * bgt_b200.h -- C ABI of libbgt_b200.so: the B200 (sm_100a) implementation of BGT's genotype hot path.
 *
 * Plain pointers and sizes only.  Every entry point names the reference interface it stands in for
 * (file:line in lh3/bgt).  The reference-side bindings (how bgt.c / pbwt.c call these) are in
 * INTEGRATION.md; the pbwt.h-compatible seam built on top of this ABI is include/pbwt_b200.h.
 *
 * Threading: one b200_ctx_t per host thread / GPU; a b200_pbf_t is immutable after load and may be
 * shared by queries of the same context.  All functions return <0 (or NULL) on error and leave a
 * message retrievable with b200_strerror().  There is no CPU fallback: without a CUDA device
 * b200_ctx_create() fails.
 */
#ifndef BGT_B200_H
#define BGT_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_MAX_GROUPS 32      /* BGT_MAX_GROUPS, bgt.h:13 */
#define B200_ABI_VERSION 2

typedef struct b200_ctx_s   b200_ctx_t;    /* one GPU: device, stream, scratch */
typedef struct b200_pbf_s   b200_pbf_t;    /* a .pbf (or a row shard of it) resident in HBM; stands in for pbf_t, pbwt.c:176-197 */
typedef struct b200_query_s b200_query_t;  /* tracked haplotypes + sample groups + site filter of one `bgt view` */

/* ---------------------------------------------------------------- context */
int         b200_abi_version(void);
int         b200_device_count(void);                 /* <=0: no usable CUDA device */
const char *b200_strerror(void);                     /* last error of the calling thread */
/* ... and its class, for callers that branch on it (the text is for people) */
enum { B200_OK = 0, B200_E_GENERIC = 1, B200_E_NO_DEVICE = 2, B200_E_CUDA = 3, B200_E_CORRUPT = 4, B200_E_FILTER_SYNTAX = 5,
       B200_E_FILTER_NEEDS_HOST = 16,   /* the filter uses `**`: evaluated with the host libm from device counts (b200_scan does it itself) */
       B200_E_UNORDERED_RECORDS = 17 }; /* b200_view_text_ex: site records are not in row order, windows cannot be mapped to row ranges */
int         b200_errcode(void);
b200_ctx_t *b200_ctx_create(int device);
void        b200_ctx_destroy(b200_ctx_t *ctx);
int         b200_ctx_sync(b200_ctx_t *ctx);
void       *b200_host_alloc(size_t bytes);           /* pinned host memory for images/results */
void        b200_host_free(void *p);
int         b200_host_register(void *p, size_t bytes);   /* pin an existing host range (a shard's byte range of a mapped .pbf) */
int         b200_host_unregister(void *p);

/* ---------------------------------------------------------------- PBF image (pbwt.c:221-262 pbf_open_r, :264-286 pbf_close) */
/* bytes = the complete .pbf file image in host memory.  Rows [row_beg,row_end) (row_end<0: to the end)
 * are made resident: the checkpoint blocks covering them are uploaded, the row offsets inside the
 * blocks are walked (the file only indexes block starts, pbwt.c:297) and the snapshots are inverted
 * into per-column start ranks.  Region sharding across GPUs = one call per rank with its own row range. */
b200_pbf_t *b200_pbf_load(b200_ctx_t *ctx, const uint8_t *bytes, size_t n_bytes, int64_t row_beg, int64_t row_end);
/* same, with B200_LOAD_* flags */
#define B200_LOAD_PREPARE_COUNT_SCAN 0x1  /* the caller will run count-only scans (`view -G`): build the composite maps of the row
                                            groups while the image is still being copied instead of at the first such scan */
b200_pbf_t *b200_pbf_load_ex(b200_ctx_t *ctx, const uint8_t *bytes, size_t n_bytes, int64_t row_beg, int64_t row_end, unsigned flags);
b200_pbf_t *b200_pbf_open(b200_ctx_t *ctx, const char *fn, int64_t row_beg, int64_t row_end);
void        b200_pbf_close(b200_pbf_t *pb);
int         b200_pbf_m(const b200_pbf_t *pb);        /* pbf_get_m, pbwt.c:391 */
int         b200_pbf_g(const b200_pbf_t *pb);        /* pbf_get_g, pbwt.c:390 */
int         b200_pbf_shift(const b200_pbf_t *pb);    /* pbf_get_shift, pbwt.c:393 */
int64_t     b200_pbf_n(const b200_pbf_t *pb);        /* pbf_get_n, pbwt.c:392 (64-bit here) */
int64_t     b200_pbf_row_beg(const b200_pbf_t *pb);  /* first resident row (multiple of 1<<shift) */
int64_t     b200_pbf_row_end(const b200_pbf_t *pb);
/* algorithmic input bytes of rows [beg,end): their 'B' records (+ the 'S' records of the checkpoints in range) */
int64_t     b200_pbf_row_bytes(const b200_pbf_t *pb, int64_t row_beg, int64_t row_end, int with_snapshots);
/* number of rows whose RLE did not sum to m (corrupt stream); the reference has undefined behaviour there */
int64_t     b200_pbf_bad_rows(const b200_pbf_t *pb);
/* resident checkpoint blocks whose count-only full-cohort scans take the split path (plane-1 carriers only) rather than the
 * general walk over every column -- decided per block on the device while loading (sparse plane 1) */
int         b200_pbf_split_blocks(const b200_pbf_t *pb);

/* ---------------------------------------------------------------- query (bgt.c:207-246 bgt_prepare, :408-416 bgtm_add_group, :444-455 bgtm_set_flt_site) */
/* out_samples: ascending sample indices (bgt_t.out, bgt.c:214-220); tracked haplotype columns are 2s and
 * 2s+1 (bgt.c:239-242).  NULL = all samples.  group: 1-based group of every selected sample (bgtm_t.group,
 * bgt.c:616); NULL = all in group 1.  flt: `-f` expression or NULL; *flt_err receives the kexpr-style parse
 * error mask (kexpr.h:10-16) and NULL is returned when it is non-zero. */
b200_query_t *b200_query_create(b200_ctx_t *ctx, const b200_pbf_t *pb, int n_out, const int32_t *out_samples,
                                const uint32_t *group, int n_groups, const char *flt, int *flt_err);
/* Column-level selection, the stand-in for pbf_subset (pbwt.c:374-388): cols = any list of column indices in any
 * order (duplicates allowed); outputs come in list order.  n_cols <= 0 or >= m or cols == NULL selects every column
 * (pbwt.c:377).  One group, no filter. */
b200_query_t *b200_query_create_cols(b200_ctx_t *ctx, const b200_pbf_t *pb, int n_cols, const int32_t *cols);
/* the same before the PBF is resident (b200_pbf_load_scan): m = columns of the file (b200_pbf_peek) */
b200_query_t *b200_query_create_m(b200_ctx_t *ctx, int m, int n_out, const int32_t *out_samples,
                                  const uint32_t *group, int n_groups, const char *flt, int *flt_err);
void          b200_query_destroy(b200_query_t *q);
int           b200_query_n_track(const b200_query_t *q);     /* 2*n_out (or the number of columns) */
int           b200_query_filter_needs_host(const b200_query_t *q); /* 1: the filter uses `**` (host libm, kexpr.c:150): b200_scan evaluates it on the host from the
                                                                * device counts; DEVICE_OUT scans and b200_view_text refuse it (B200_E_FILTER_NEEDS_HOST) */
int           b200_query_hap_words(const b200_query_t *q);   /* 32-bit words per plane per row of hap_bits */
int           b200_query_counts_stride(const b200_query_t *q); /* 3 + 3*n_groups */

/* ---------------------------------------------------------------- the scan (pbwt.c:313-337 pbf_read per row + bgt.c:735-757 bgtm_cal_info + bgt.c:712-719 bgtm_pass_site_flt) */
#define B200_SCAN_COUNTS     0x01  /* per-site AN/AC[/AN#/AC#] */
#define B200_SCAN_HAP_BITS   0x02  /* genotype rows as two bit planes, bit i of word w = tracked haplotype 32w+i */
#define B200_SCAN_HAP_BYTES  0x04  /* genotype rows as pbf_read returns them: one byte per tracked haplotype per plane */
#define B200_SCAN_DEVICE_OUT 0x10  /* output pointers are device pointers; no D2H, call returns without syncing */
#define B200_SCAN_NO_SPLIT   0x20  /* testing: always walk every tracked column (disable the split scan of count-only full-cohort queries) */
#define B200_SCAN_NO_COMPOSE 0x40  /* testing: split scan without the composite maps of row groups */
#define B200_SCAN_NO_SEGMENTS 0x80 /* testing: per-group marginals of the split scan with one CTA per checkpoint block (no segment vectors) */
#define B200_SCAN_NO_PIECES  0x1000 /* testing: per-group marginals of the split scan by the row loop over bit vectors (marginal.cu) instead of the piece lists (margpiece.cu) */
#define B200_SCAN_COLS_PER_THREAD(c) ((unsigned)(c) << 8)  /* testing/tuning: force 1, 2, 4 or 8 tracked columns per thread (0 = automatic) */

typedef struct {
	int32_t  *counts;        /* [n_rows][3+3G]: AN, AC(first ALT), AC(<M>), then (AN#, AC#, AC#<M>) per group (bgt_info_t, bgt.h:44-47) */
	uint8_t  *pass;          /* [n_rows]: 1 if the site passes the filter (always 1 without a filter) */
	uint32_t *hap_bits[2];   /* [n_rows][hap_words] per plane */
	uint8_t  *hap_bytes[2];  /* [n_rows][n_track] per plane */
	int64_t   totals[4];     /* out: sum AN, sum AC, sum AC<M>, rows passed -- the per-shard figures the multi-GPU all-reduce sums */
} b200_scan_out_t;

/* Scan rows [row_beg,row_beg+n_rows) (must be resident).  Unwanted outputs: NULL pointers / flags off.
 * Returns rows scanned or <0. */
int64_t b200_scan(b200_ctx_t *ctx, const b200_pbf_t *pb, const b200_query_t *q, int64_t row_beg, int64_t n_rows,
                  unsigned flags, b200_scan_out_t *out);

/* header fields of a .pbf image (pbwt.c:231-258): columns, snapshot interval, rows; host logic only */
int     b200_pbf_peek(const uint8_t *bytes, size_t n_bytes, int32_t *m, int32_t *shift, int64_t *n_rows);
/* b200_pbf_load_ex(B200_LOAD_PREPARE_COUNT_SCAN) + b200_scan(B200_SCAN_COUNTS) of rows [row_beg,row_end) as ONE pipeline: for the
 * count-only full-cohort scan with one group (`view -f .. -G`) the pair walk, the per-site AC/AN + verdict and the copy of the
 * results to out->counts / out->pass (host memory, pinned for speed) are queued behind the composite maps of every chunk of
 * the image while later chunks are still being copied; other queries get the load followed by the scan.  Returns the
 * resident handle (close it with b200_pbf_close) or NULL; *n_scanned = rows produced. */
b200_pbf_t *b200_pbf_load_scan(b200_ctx_t *ctx, const uint8_t *bytes, size_t n_bytes, int64_t row_beg, int64_t row_end,
                               const b200_query_t *q, b200_scan_out_t *out, int64_t *n_scanned);
/* Batched regions (`-r` / `-B` / server-style short queries; in the reference one pbf_seek + pbf_read loop per region,
 * pbwt.c:349-372): the scans of all regions are queued back to back, their results come home with one copy and one
 * synchronisation.  Regions must be resident; rows of region i land at index sum(n_rows[0..i)) of every output; totals are
 * summed over the regions.  Returns the rows produced or <0. */
int64_t b200_scan_regions(b200_ctx_t *ctx, const b200_pbf_t *pb, const b200_query_t *q, int n_regions, const int64_t *row_beg,
                          const int64_t *n_rows, unsigned flags, b200_scan_out_t *out);
/* After a B200_SCAN_DEVICE_OUT scan: wait for it, fetch its totals and kernel timings. */
int     b200_scan_collect(b200_ctx_t *ctx, int64_t totals[4]);
/* totals (as b200_scan_out_t.totals) of the last scan of this context whose results reached the host: b200_scan,
 * b200_scan_collect, or the scan inside b200_view_text[_ex] */
int     b200_last_totals(b200_ctx_t *ctx, int64_t totals[4]);

/* ---------------------------------------------------------------- multi-GPU (view.c:150-155 is the loop that gets sharded) */
/* Region shards = row ranges of whole checkpoint blocks (pbwt.c:292-301 snapshots, :268-276 block index): every GPU of a
 * process gets its own context and b200_pbf_load[_ex](row_beg,row_end); no per-site exchange exists.  The one collective
 * sums the per-shard totals: vals[n][count] int64, in = each context's values, out = the sums in every row
 * (ncclAllReduce(sum) over the contexts' GPUs, one communicator per call; n == 1 is a no-op).  NCCL (libnccl.so.2) is
 * loaded when this is first called, not linked.  Processes that own one GPU each (bench.py under torchrun) use their
 * launcher's communicator instead. */
int     b200_allreduce_i64(b200_ctx_t *const *ctxs, int n, int64_t *vals, int count);

/* device time (ms) of the kernels of the last b200_scan on this context, measured with CUDA events on the
 * context's stream: which = 0 all decode phases (plane-1 select + rank walk + group marginals), 1 whole scan (all
 * kernels), 2 H2D of the last load, 3 D2H of the results, 4 plane1_select_kernel, 5 pbwt_marginal_kernel.
 * The rank-walk kernel alone is [0] - [4] - [5]. */
double  b200_last_ms(b200_ctx_t *ctx, int which);
int64_t b200_kernel_launches(b200_ctx_t *ctx);      /* kernels launched by this context so far */
/* user timing marks on the context's stream (slot 0..3): record now / device ms between two recorded marks (syncs) */
int     b200_mark(b200_ctx_t *ctx, int slot);
double  b200_mark_elapsed_ms(b200_ctx_t *ctx, int slot_a, int slot_b);

/* ---------------------------------------------------------------- host logic, callable without a device */
/* Compile `flt` exactly as b200_query_create does and evaluate it on the host for n_rows count vectors
 * ([n_rows][3+3*n_groups], the layout b200_scan writes).  Returns the kexpr-style parse error mask (0 = ok).
 * This is the evaluator the scan itself uses for filters containing `**` (host libm). */
int     b200_flt_eval_host(const char *flt, int n_groups, const int32_t *counts, int64_t n_rows, uint8_t *pass);
/* Parse a .pbf image, walk the row index of rows [row_beg,row_end) and plan the kernel's row tiles.
 * info[0..7] = m, shift, rows in file, resident blocks, tiles, oversized single-row tiles, largest ordinary tile in
 * bytes, largest row in bytes.  Returns 0 or <0 (message in b200_strerror()). */
int     b200_pbf_plan(const uint8_t *bytes, size_t n_bytes, int64_t row_beg, int64_t row_end, int64_t info[8]);

/* ---------------------------------------------------------------- synthetic cohort (SURVEY 8d generator; rows drawn in PBWT-rank space, truthful snapshots) */
typedef struct {
	int32_t  n_samples;      /* m = 2*n_samples */
	int64_t  n_rows;
	int32_t  shift;          /* 13 */
	uint64_t seed;
	int32_t  r_max;          /* plane 0: the row's allele count is log-uniform in [1,m/2], spread over 1+U[0,r_max) intervals of 1s in rank order */
	int32_t  p1_one_in;      /* plane 1 (missing / other-ALT codes) non-empty in one of p1_one_in rows (16) ... */
	int32_t  p1_max_iv;      /* ... with 1 + U[0,p1_max_iv) intervals (0 = default 3) ... */
	int32_t  p1_max_len;     /* ... of 1 + U[0,p1_max_len) ones each (0 = default 64) */
} b200_synth_t;
/* Generates the .pbf image on the device and returns it resident for rows [0,n_rows). */
b200_pbf_t *b200_synth_generate(b200_ctx_t *ctx, const b200_synth_t *cfg);
size_t      b200_pbf_image_size(const b200_pbf_t *pb);             /* bytes of the complete file image, 0 if only a shard is held */
int         b200_pbf_image_download(const b200_pbf_t *pb, uint8_t *dst, size_t n_bytes); /* device image -> host */
int         b200_pbf_image_download_range(const b200_pbf_t *pb, uint8_t *dst, uint64_t off, size_t n_bytes);
/* file byte range of the checkpoint blocks holding rows [row_beg,row_end) and the offset of the index record: a region
 * shard's b200_pbf_load_ex touches the 16-byte header, [byte_beg,byte_end) and [index_beg, end of file) of the image only */
int         b200_pbf_block_bytes(const b200_pbf_t *pb, int64_t row_beg, int64_t row_end, uint64_t *byte_beg, uint64_t *byte_end, uint64_t *index_beg);

/* ---------------------------------------------------------------- encoder (pbwt.c:199-219 pbf_open_w, :288-311 pbf_write, :264-286 pbf_close) */
/* The PBWT encoder on the device: the column-owned rank walk run forward (pbc_enc_core, pbwt.c:57-66) + the run-length
 * byte code (pbr_enc, pbwt.c:24-50).  g = 2 bit planes.  The .pbf image is assembled in host memory like pbf_write
 * assembles the file: header, an 'S' snapshot before every 2^shift rows, one 'B' record per row, the index at finish. */
typedef struct b200_enc_s b200_enc_t;
b200_enc_t *b200_enc_create(b200_ctx_t *ctx, int m, int shift);
/* rows as pbf_write takes them (uint8_t *const *a, pbwt.c:288): a0/a1 = [n_rows][m], one byte per haplotype per plane, non-zero = 1 (pbwt.c:61) */
int         b200_enc_write_bytes(b200_enc_t *e, const uint8_t *a0, const uint8_t *a1, int64_t n_rows);
/* the same rows as bit planes in column order: bits[row][plane][(m+31)/32], bit c of word w = haplotype 32w+c */
int         b200_enc_write_bits(b200_enc_t *e, const uint32_t *bits, int64_t n_rows);
int64_t     b200_enc_rows(const b200_enc_t *e);
/* streaming (pbf_write writes each row as it goes, pbwt.c:288-311): *bytes = the file bytes assembled since the last drain
 * (owned by the encoder, valid until the next call on it); they are forgotten by the encoder.  Returns their number. */
int64_t     b200_enc_drain(b200_enc_t *e, const uint8_t **bytes);
/* writes the index record (pbf_close, pbwt.c:268-276); *image = the complete file image -- or, after b200_enc_drain calls, the
 * rest of it behind the drained bytes -- (owned by the encoder, valid until b200_enc_destroy); returns its size */
int64_t     b200_enc_finish(b200_enc_t *e, const uint8_t **image);
void        b200_enc_destroy(b200_enc_t *e);

/* ---------------------------------------------------------------- BGZF (bgzf.c:225-249 inflate_block, :318-351 bgzf_read_block, :353-379 bgzf_read) */
/* Inflate a whole BGZF file image (the site-only .bcf / .csi of a BGT database) on the device: one warp per block, raw
 * DEFLATE (RFC 1951), the CRC is not checked (like the reference's reader).  out == NULL: returns the uncompressed size
 * (block headers only).  Returns the number of bytes written, <0 on a malformed file. */
int64_t     b200_bgzf_inflate(b200_ctx_t *ctx, const uint8_t *bytes, size_t n_bytes, uint8_t *out, size_t out_cap);

/* ---------------------------------------------------------------- sites + the text of `bgt view -G` (output assembly and BCF parse off the host thread) */
/* The site side of a BGT database on the device: <prefix>.bgt.bcf (site-only BCF2, bgzf) and <prefix>.bgt.bcf.csi.  Both are
 * inflated by the kernel above; the record-number index of the .csi (RNI, hts.c:536-542: one virtual offset per 1024
 * records) gives independent starting points for the record-length chase (vcf.c:316-336 bcf_read1_core); every record is
 * parsed by one thread (vcf.c:338-360 bcf_unpack; INFO/_row, bgt.c:279-286).  csi may be NULL (one sequential chase).
 * row_key: header-dictionary id of INFO/_row (bcf_id2int(h, BCF_DT_ID, "_row")), or -1 = take it from the header text. */
typedef struct b200_sites_s b200_sites_t;
b200_sites_t *b200_sites_load(b200_ctx_t *ctx, const uint8_t *bcf, size_t n_bcf, const uint8_t *csi, size_t n_csi, int row_key);
int64_t     b200_sites_n(const b200_sites_t *s);                        /* records */
int         b200_sites_rows_sorted(const b200_sites_t *s);              /* 1: INFO/_row ascends with the records (what `bgt import` writes) */
/* the records whose row lies in [row_beg,row_end) (records in row order): maps a region shard to its record window */
int         b200_sites_rec_range(const b200_sites_t *s, int64_t row_beg, int64_t row_end, int64_t *rec_beg, int64_t *rec_end);
const char *b200_sites_header(const b200_sites_t *s, int64_t *len);      /* BCF header text (vcf.c:263-288) */
int         b200_sites_rows(const b200_sites_t *s, int64_t *rows, int32_t *pos);  /* per record: INFO/_row and 0-based POS */
void        b200_sites_destroy(b200_sites_t *s);
/* `bgt view -G [-C] [-f EXPR] [-s ...]` over the whole resident PBF: the scan (b200_scan) with its results kept on the
 * device, then one VCF line per site that passes, assembled on the device in file order -- byte for byte what
 * vcf_format1 (vcf.c:895-969) prints for the record that bgtm_read_core builds (bcfcpy_min vcf.c:1166-1182, END bgt.c:824-827,
 * bgtm_fill_info bgt.c:721-733).  with_counts = the -C flag (forced on by a filter or several groups, bgt.c:850).
 * contig_names: the output header's contig dictionary, or NULL = the .bcf's own.  *text = pinned host buffer owned by
 * `s`, valid until the next call; returns its length, *n_lines = records printed.  The header lines are the caller's. */
int64_t     b200_view_text(b200_ctx_t *ctx, b200_sites_t *s, const b200_pbf_t *pb, const b200_query_t *q, int with_counts,
                           const char *const *contig_names, int n_contigs, const char **text, int64_t *n_lines);
/* The same for records [rec_beg, rec_end) only (rec_end < 0: to the last), optionally WITH the genotype columns
 * (`bgt view` without -G, e.g. a 200-sample subset: FORMAT GT and one unphased `a/b` per selected sample from the
 * scan's bit planes, bgt_gen_gt bgt.c:290-313 + vcf.c:940-966).  Only the rows behind the window are scanned, so a
 * caller bounds the text per call by the window (records must be in row order for windows smaller than the file). */
#define B200_VIEW_COUNTS    0x1   /* -C */
#define B200_VIEW_GENOTYPES 0x2   /* no -G */
int64_t     b200_view_text_ex(b200_ctx_t *ctx, b200_sites_t *s, const b200_pbf_t *pb, const b200_query_t *q, unsigned flags,
                              int64_t rec_beg, int64_t rec_end, const char *const *contig_names, int n_contigs,
                              const char **text, int64_t *n_lines);

#ifdef __cplusplus
}
#endif
#endif
