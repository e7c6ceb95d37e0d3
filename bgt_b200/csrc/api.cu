// api.cu -- the C ABI of libbgt_b200.so (include/bgt_b200.h): handles, HBM residency, launch orchestration.
// Host logic only; every per-site computation is in pbwt_kernels.cu / synth.cu.  No CPU fallback exists:
// without a CUDA device b200_ctx_create() fails and nothing else can be called.
#include <cuda_runtime.h>
#include <nccl.h>      // types and prototypes only: libnccl.so.2 is loaded on the first b200_allreduce_i64 call, not linked
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <time.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <mutex>
#include <vector>
#include <algorithm>
#include <string>
#include "../../include/bgt_b200.h"
#include "pbwt_kernels.cuh"
#include "flt.h"

using namespace b200;

// ------------------------------------------------------------------------------------------------ errors

static thread_local char g_err[512] = "";
static thread_local int g_errcode = B200_OK;

static void set_err(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	g_errcode = B200_E_GENERIC;
}

// same, with one of the B200_E_* codes a caller can branch on (b200_errcode) instead of reading the text
static void set_err_code(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	g_errcode = code;
}

static double now_ms()
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

#define CU_OK(call) cu_ok((call), #call, __LINE__)
static bool cu_ok(cudaError_t e, const char *what, int line)
{
	if (e == cudaSuccess) return true;
	set_err_code(B200_E_CUDA, "CUDA error at api.cu:%d: %s: %s", line, what, cudaGetErrorString(e));
	return false;
}

// ------------------------------------------------------------------------------------------------ handles

struct DevBuf { // grow-only device scratch
	void *p = nullptr; size_t cap = 0;
	bool reserve(size_t n) {
		if (n <= cap) return true;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		if (!CU_OK(cudaMalloc(&p, n))) return false;
		cap = n;
		return true;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PoolBlock { void *p; size_t size; };
constexpr int N_IDX_STREAMS = 8;
constexpr int LOAD_CHUNKS = 16;   // a PBF image is copied in up to this many chunks of whole checkpoint blocks

struct b200_ctx_s {
	// device-memory pool: PBF windows are loaded and dropped repeatedly (seam A/B, end-to-end scans); cudaMalloc /
	// cudaFree of hundreds of MB cost milliseconds to hundreds of milliseconds, so freed blocks are kept for reuse
	std::vector<PoolBlock> pool_free_list, pool_live;
	size_t pool_cached = 0, pool_limit = (size_t)16 << 30;   // cached bytes are capped at half of the device memory
	int dev = 0;
	cudaStream_t st = nullptr;
	cudaStream_t st_copy = nullptr;   // H2D of PBF images, so that per-chunk kernels on `st` overlap the rest of the copy
	// fused load + scan: the pair walk + finalize of a chunk run on st_walk beside the composite maps of the next chunks (st), and
	// the chunk's results go home on st_d2h
	cudaStream_t st_d2h = nullptr, st_walk = nullptr;
	// the composite maps of consecutive chunks alternate between st and st_comp2: the last wave of one launch overlaps the first of the next
	cudaStream_t st_comp2 = nullptr;
	cudaEvent_t ev_comp2 = nullptr;
	cudaEvent_t ev_fin[LOAD_CHUNKS + 1] = {}, ev_comp[LOAD_CHUNKS + 1] = {};
	bool fused_pairs_done = false;    // b200_scan called from the fused path: cnt_raw already holds the pair-walk counts
	bool keep_totals = false;         // b200_scan called for a later region of b200_scan_regions: totals and error flags accumulate
	DevBuf rg_counts, rg_pass, rg_bits[2], rg_bytes[2];   // outputs of a batch of regions
	cudaStream_t st_idx[N_IDX_STREAMS] = {};      // the row-index chase of a chunk is one long dependent chain per block (latency bound, a few
	                                  // lanes): the chases of consecutive chunks overlap each other and the kernels of earlier chunks
	cudaEvent_t ev_chunk[LOAD_CHUNKS] = {}, ev_idx[LOAD_CHUNKS] = {}, ev_sel[LOAD_CHUNKS] = {};
	cudaEvent_t ev[12] = {};  // 0/1 walk phase(s), 2/3 scan, 4/5 h2d, 6/7 d2h, 8/9 plane1 select, 10/11 marginals
	cudaEvent_t mark[4] = {};
	cudaEvent_t ev_zero = nullptr;
	double last_ms[6] = {0, 0, 0, 0, 0, 0};
	int64_t last_totals[4] = {0, 0, 0, 0};   // totals of the last scan whose results reached the host (b200_last_totals)
	int64_t launches = 0;
	// device error flags, one word per kind of operation so that they cannot wipe or taint each other:
	// d_err = loads (row index, snapshots, composites built while loading), d_err_scan = b200_scan (zeroed when a scan
	// starts, reported when it ends or is collected), d_err_sites = inflate / BCF parse / text assembly
	int *d_err = nullptr, *d_err_scan = nullptr, *d_err_sites = nullptr;
	unsigned long long *d_acc = nullptr; // [0..3] totals, [4] bad rows
	DevBuf cnt_raw, counts, pass, hapbits[2], hapbytes[2], qcol, qrow, qcount, n0g, vseg, seg_ok;
	int sm_count = 148;
	bool split_used = false, marginal_used = false;
};

struct b200_pbf_s {
	b200_ctx_t *ctx = nullptr;
	std::vector<uint64_t> h_idx;      // the file's block index (pbwt.c:268-276)
	std::vector<int> chunk_blk;       // load in flight: resident-block boundaries of the H2D chunks (events ctx->ev_chunk[k])
	bool prepare_split = false;       // build the composite maps while loading (b200_pbf_load_ex)
	uint64_t ioff = 0;                // file offset of the 'I' record
	int m = 0, g = 0, shift = 0, BS = 0;
	int64_t n = 0;               // rows in the file
	int blk0 = 0, n_blk = 0;     // resident checkpoint blocks [blk0, blk0+n_blk)
	int64_t n_blk_file = 0;
	std::vector<int> rows_in_blk;
	std::vector<uint64_t> h_rowoff;   // [n_blk][BS+1], relative to d_img: host copy of d_rowoff, fetched on demand (host_rowoff())
	std::vector<uint64_t> h_blkoff;   // [n_blk] offset of the 'S' record, relative to d_img
	std::vector<uint64_t> h_blkend;   // [n_blk] end of the block's records (next 'S' record or the 'I' record)
	uint8_t *d_img = nullptr; size_t img_bytes = 0; uint64_t file_off0 = 0;
	size_t file_size = 0;             // size of the complete file image (0 unless fully resident)
	uint64_t *d_rowoff = nullptr, *d_blkoff = nullptr, *d_blkend = nullptr, *d_ix_scratch = nullptr;
	int *d_rows_in_blk = nullptr, *d_blk_tile_beg = nullptr, *d_blk_tile_end = nullptr;
	uint8_t *d_blk_sparse = nullptr;
	int2 *d_tiles = nullptr;
	uint32_t *d_n1 = nullptr, *d_nrun0 = nullptr;
	int32_t *d_rank0 = nullptr;
	int64_t bad_rows = 0;
	// "plane-1 view" of the resident blocks: only the rows whose second bit plane (missing / other-ALT codes) is not
	// empty, as a miniature PBF image the walk kernel can run on (first phase of the split scan)
	bool p1_ready = false;
	int p1_cap = 0;                      // capacity of a block's (column,row) pair list
	int p1_base = 0;                     // pairs the launches queued without knowing the data cover (the load pipeline's); blocks flagged 2 hold more
	std::vector<uint8_t> blk_sparse;     // [n_blk] != 0: the block's plane-1 ones fit p1_cap (split scan applies); 2: more than p1_base of them
	std::vector<uint32_t> blk_ones;      // [n_blk] plane-1 ones (= pairs) of every block
	uint32_t *d_blk_ones = nullptr;
	bool sel_ext_ready = false;          // the select has been extended to the slices behind p1_base of the blocks flagged 2: their pairs
	int32_t *d_qcol_ext = nullptr; uint16_t *d_qrow_ext = nullptr; long long *d_ext_off = nullptr;   // from p1_base on live here, block after block
	std::vector<long long> ext_off;
	uint8_t *d_p1img = nullptr;
	uint64_t *d_p1_rowoff = nullptr;
	uint32_t *d_p1_n1 = nullptr, *d_p1_prefix = nullptr;
	uint16_t *d_p1_realrow = nullptr;
	int *d_grp_tile_beg = nullptr;       // [n_blk][groups+1] first tile of every COMP_K-row group
	// composite plane-0 maps of the row groups (compose.cu), built lazily by the first split scan and cached
	mutable bool comp_ready = false;
	mutable uint32_t *d_comp_start = nullptr;
	mutable int32_t *d_comp_delta = nullptr;
	mutable int *d_comp_n = nullptr;
	mutable uint16_t *d_comp_dir = nullptr;   // bucket directories of the forward composites
	// the (column, row) pairs that carry a plane-1 bit, per sparse block (plane1_select_kernel): query independent, so they
	// are found once per resident PBF together with the composites
	mutable int32_t *d_qcol = nullptr; mutable uint16_t *d_qrow = nullptr; mutable int *d_qcount = nullptr;
	mutable bool sel_ready = false;
	int dir_shift = 0, dir_n = 0;
	mutable uint32_t *d_vcomp_start = nullptr;   // inverse composites of the plane-1 view rows (plane1_select_kernel)
	mutable int32_t *d_vcomp_delta = nullptr;
	mutable int *d_vcomp_n = nullptr;
	mutable uint16_t *d_vcomp_dir = nullptr;
	int *d_p1_rows_in_blk = nullptr;
	long long *d_p1_vbase = nullptr;     // [n_blk+1] first view row of every block
	int64_t p1_rows = 0;
};

struct b200_query_s {
	b200_ctx_t *ctx = nullptr;
	int m = 0, n_out = 0, n_track = 0, G = 1, words = 0;
	bool full = true;
	int32_t *d_track = nullptr;
	uint8_t *d_tgrp = nullptr;
	int32_t *d_gsize = nullptr;
	flt_prog_t prog;
	flt_prog_t *d_prog = nullptr;
	bool has_flt = false;
};

// ------------------------------------------------------------------------------------------------ device-memory pool

static bool pool_malloc(b200_ctx_t *c, void **out, size_t bytes)
{
	if (bytes == 0) bytes = 16;
	int best = -1;
	for (size_t i = 0; i < c->pool_free_list.size(); ++i) {
		const size_t sz = c->pool_free_list[i].size;
		if (sz >= bytes && sz <= bytes * 2 + (1u << 20) && (best < 0 || sz < c->pool_free_list[best].size)) best = (int)i;
	}
	if (best >= 0) {
		PoolBlock b = c->pool_free_list[best];
		c->pool_free_list.erase(c->pool_free_list.begin() + best);
		c->pool_cached -= b.size;
		c->pool_live.push_back(b);
		*out = b.p;
		return true;
	}
	void *p = nullptr;
	if (cudaMalloc(&p, bytes) != cudaSuccess) { // out of memory: drop the cache and retry once
		cudaGetLastError();
		for (auto &b : c->pool_free_list) cudaFree(b.p);
		c->pool_free_list.clear(); c->pool_cached = 0;
		if (!CU_OK(cudaMalloc(&p, bytes))) return false;
	}
	c->pool_live.push_back({p, bytes});
	*out = p;
	return true;
}

static void pool_free(b200_ctx_t *c, void *p)
{
	if (!p) return;
	for (size_t i = 0; i < c->pool_live.size(); ++i)
		if (c->pool_live[i].p == p) {
			PoolBlock b = c->pool_live[i];
			c->pool_live.erase(c->pool_live.begin() + i);
			if (c->pool_cached + b.size > c->pool_limit) { cudaFree(b.p); return; }
			c->pool_free_list.push_back(b);
			c->pool_cached += b.size;
			return;
		}
	cudaFree(p);
}

// ------------------------------------------------------------------------------------------------ context

extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }

extern "C" int b200_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

extern "C" const char *b200_strerror(void) { return g_err; }
extern "C" int b200_errcode(void) { return g_errcode; }

extern "C" b200_ctx_t *b200_ctx_create(int device)
{
	const bool trace = getenv("BGT_B200_TRACE") != nullptr;
	const double t0 = now_ms();
	int n = b200_device_count();
	const double t1 = now_ms();
	if (n <= 0) { set_err_code(B200_E_NO_DEVICE, "no CUDA device: libbgt_b200 has no CPU fallback"); return nullptr; }
	if (device < 0 || device >= n) { set_err("device %d out of range (0..%d)", device, n - 1); return nullptr; }
	if (!CU_OK(cudaSetDevice(device)) || !CU_OK(cudaFree(nullptr))) return nullptr;
	const double t2 = now_ms();
	cudaDeviceProp prop;
	if (!CU_OK(cudaGetDeviceProperties(&prop, device))) return nullptr;
	if (prop.major < 10) { set_err("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor); return nullptr; }
	b200_ctx_t *c = new b200_ctx_t();
	c->dev = device; c->sm_count = prop.multiProcessorCount;
	c->pool_limit = prop.totalGlobalMem / 2;
	bool ok = CU_OK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking)) && CU_OK(cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking)) &&
	          CU_OK(cudaStreamCreateWithFlags(&c->st_d2h, cudaStreamNonBlocking)) && CU_OK(cudaStreamCreateWithFlags(&c->st_comp2, cudaStreamNonBlocking)) &&
	          CU_OK(cudaEventCreateWithFlags(&c->ev_comp2, cudaEventDisableTiming));
	for (int i = 0; ok && i <= LOAD_CHUNKS; ++i) ok = CU_OK(cudaEventCreateWithFlags(&c->ev_fin[i], cudaEventDisableTiming)) && CU_OK(cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
	{ // the small latency-bound index kernels must not queue behind the wide kernels they overlap: highest priority
		int prio_lo = 0, prio_hi = 0;
		cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
		for (int i = 0; ok && i < N_IDX_STREAMS; ++i) ok = CU_OK(cudaStreamCreateWithPriority(&c->st_idx[i], cudaStreamNonBlocking, prio_hi));
		// the walks of the fused pipeline follow their chunk's maps closely instead of queueing up behind the maps of later chunks
		ok = ok && CU_OK(cudaStreamCreateWithPriority(&c->st_walk, cudaStreamNonBlocking, prio_hi));
	}
	for (int i = 0; ok && i < LOAD_CHUNKS; ++i) ok = CU_OK(cudaEventCreateWithFlags(&c->ev_chunk[i], cudaEventDisableTiming)) && CU_OK(cudaEventCreateWithFlags(&c->ev_idx[i], cudaEventDisableTiming)) &&
	                                              CU_OK(cudaEventCreateWithFlags(&c->ev_sel[i], cudaEventDisableTiming));
	for (int i = 0; ok && i < 12; ++i) ok = CU_OK(cudaEventCreate(&c->ev[i]));
	for (int i = 0; ok && i < 4; ++i) ok = CU_OK(cudaEventCreate(&c->mark[i]));
	ok = ok && CU_OK(cudaEventCreateWithFlags(&c->ev_zero, cudaEventDisableTiming));
	ok = ok && CU_OK(cudaMalloc(&c->d_err, 4 * sizeof(int))) && CU_OK(cudaMalloc(&c->d_acc, 8 * sizeof(unsigned long long)));
	ok = ok && CU_OK(cudaMemset(c->d_err, 0, 4 * sizeof(int))) && CU_OK(cudaMemset(c->d_acc, 0, 8 * sizeof(unsigned long long)));
	if (!ok) { b200_ctx_destroy(c); return nullptr; }
	c->d_err_scan = c->d_err + 1; c->d_err_sites = c->d_err + 2;
	if (trace) fprintf(stderr, "[b200 trace] context: driver init + device count %.1f ms, primary context %.1f ms, streams/events/scratch %.1f ms\n", t1 - t0, t2 - t1, now_ms() - t2);
	return c;
}

extern "C" void b200_ctx_destroy(b200_ctx_t *c)
{
	if (!c) return;
	cudaSetDevice(c->dev);
	if (c->st) cudaStreamSynchronize(c->st);
	if (c->st_copy) cudaStreamSynchronize(c->st_copy);
	for (int i = 0; i < N_IDX_STREAMS; ++i) if (c->st_idx[i]) cudaStreamSynchronize(c->st_idx[i]);
	c->cnt_raw.release(); c->counts.release(); c->pass.release();
	for (int p = 0; p < 2; ++p) { c->hapbits[p].release(); c->hapbytes[p].release(); }
	c->rg_counts.release(); c->rg_pass.release(); for (int p = 0; p < 2; ++p) { c->rg_bits[p].release(); c->rg_bytes[p].release(); }
	c->n0g.release(); c->vseg.release(); c->seg_ok.release(); c->qcol.release(); c->qrow.release(); c->qcount.release();
	for (int i = 0; i < 12; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
	for (int i = 0; i < 4; ++i) if (c->mark[i]) cudaEventDestroy(c->mark[i]);
	if (c->ev_zero) cudaEventDestroy(c->ev_zero);
	for (auto &b : c->pool_free_list) cudaFree(b.p);
	for (auto &b : c->pool_live) cudaFree(b.p);
	if (c->d_err) cudaFree(c->d_err);
	if (c->d_acc) cudaFree(c->d_acc);
	for (int i = 0; i < LOAD_CHUNKS; ++i) { if (c->ev_chunk[i]) cudaEventDestroy(c->ev_chunk[i]); if (c->ev_idx[i]) cudaEventDestroy(c->ev_idx[i]); if (c->ev_sel[i]) cudaEventDestroy(c->ev_sel[i]); }
	if (c->st_copy) cudaStreamDestroy(c->st_copy);
	if (c->st_walk) { cudaStreamSynchronize(c->st_walk); cudaStreamDestroy(c->st_walk); }
	if (c->st_d2h) { cudaStreamSynchronize(c->st_d2h); cudaStreamDestroy(c->st_d2h); }
	if (c->st_comp2) { cudaStreamSynchronize(c->st_comp2); cudaStreamDestroy(c->st_comp2); }
	if (c->ev_comp2) cudaEventDestroy(c->ev_comp2);
	for (int i = 0; i <= LOAD_CHUNKS; ++i) { if (c->ev_fin[i]) cudaEventDestroy(c->ev_fin[i]); if (c->ev_comp[i]) cudaEventDestroy(c->ev_comp[i]); }
	for (int i = 0; i < N_IDX_STREAMS; ++i) if (c->st_idx[i]) cudaStreamDestroy(c->st_idx[i]);
	if (c->st) cudaStreamDestroy(c->st);
	delete c;
}

extern "C" int b200_ctx_sync(b200_ctx_t *c)
{
	if (!c) return -1;
	cudaSetDevice(c->dev);
	return CU_OK(cudaStreamSynchronize(c->st)) ? 0 : -1;
}

extern "C" void *b200_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (!CU_OK(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault))) return nullptr;
	return p;
}

extern "C" void b200_host_free(void *p) { if (p) cudaFreeHost(p); }

// pin an existing host range (e.g. the part of a mapped .pbf a region shard will upload) so that its H2D copy runs at DMA speed
extern "C" int b200_host_register(void *p, size_t bytes)
{
	if (!p || !bytes) { set_err("b200_host_register: null range"); return -1; }
	return CU_OK(cudaHostRegister(p, bytes, cudaHostRegisterDefault)) ? 0 : -1;
}
extern "C" int b200_host_unregister(void *p) { return p && CU_OK(cudaHostUnregister(p)) ? 0 : -1; }

extern "C" double b200_last_ms(b200_ctx_t *c, int which) { return (c && which >= 0 && which < 6) ? c->last_ms[which] : -1.0; }
extern "C" int64_t b200_kernel_launches(b200_ctx_t *c) { return c ? c->launches : 0; }

extern "C" int b200_mark(b200_ctx_t *c, int slot)
{
	if (!c || slot < 0 || slot >= 4) return -1;
	cudaSetDevice(c->dev);
	return CU_OK(cudaEventRecord(c->mark[slot], c->st)) ? 0 : -1;
}

extern "C" double b200_mark_elapsed_ms(b200_ctx_t *c, int a, int b)
{
	if (!c || a < 0 || a >= 4 || b < 0 || b >= 4) return -1.0;
	cudaSetDevice(c->dev);
	float ms = 0;
	if (!CU_OK(cudaEventSynchronize(c->mark[b])) || !CU_OK(cudaEventElapsedTime(&ms, c->mark[a], c->mark[b]))) return -1.0;
	return ms;
}

// ------------------------------------------------------------------------------------------------ PBF residency

static void pbf_free_device(b200_pbf_t *pb)
{
	pool_free(pb->ctx, pb->d_img);
	pool_free(pb->ctx, pb->d_rowoff);
	pool_free(pb->ctx, pb->d_blkoff);
	pool_free(pb->ctx, pb->d_rows_in_blk);
	pool_free(pb->ctx, pb->d_blk_tile_beg);
	pool_free(pb->ctx, pb->d_blk_tile_end);
	pool_free(pb->ctx, pb->d_blkend);
	pool_free(pb->ctx, pb->d_ix_scratch);
	pool_free(pb->ctx, pb->d_blk_sparse); pool_free(pb->ctx, pb->d_blk_ones);
	pool_free(pb->ctx, pb->d_qcol_ext); pool_free(pb->ctx, pb->d_qrow_ext); pool_free(pb->ctx, pb->d_ext_off);
	pool_free(pb->ctx, pb->d_tiles);
	pool_free(pb->ctx, pb->d_n1);
	pool_free(pb->ctx, pb->d_nrun0);
	pool_free(pb->ctx, pb->d_rank0);
	pool_free(pb->ctx, pb->d_p1img);
	pool_free(pb->ctx, pb->d_p1_rowoff);
	pool_free(pb->ctx, pb->d_p1_n1);
	pool_free(pb->ctx, pb->d_p1_prefix);
	pool_free(pb->ctx, pb->d_p1_realrow);
	pool_free(pb->ctx, pb->d_grp_tile_beg);
	pool_free(pb->ctx, pb->d_comp_start);
	pool_free(pb->ctx, pb->d_comp_delta);
	pool_free(pb->ctx, pb->d_comp_n);
	pool_free(pb->ctx, pb->d_comp_dir);
	pool_free(pb->ctx, pb->d_qcol); pool_free(pb->ctx, pb->d_qrow); pool_free(pb->ctx, pb->d_qcount);
	pool_free(pb->ctx, pb->d_vcomp_start);
	pool_free(pb->ctx, pb->d_vcomp_delta);
	pool_free(pb->ctx, pb->d_vcomp_n);
	pool_free(pb->ctx, pb->d_vcomp_dir);
	pool_free(pb->ctx, pb->d_p1_rows_in_blk);
	pool_free(pb->ctx, pb->d_p1_vbase);
}

extern "C" void b200_pbf_close(b200_pbf_t *pb)
{
	if (!pb) return;
	cudaSetDevice(pb->ctx->dev);
	cudaStreamSynchronize(pb->ctx->st);
	pbf_free_device(pb);
	delete pb;
}

// Tiles: maximal groups of consecutive rows of one block with <= T_MAX rows and <= RAW_CAP bytes; a row that is
// larger than RAW_CAP on its own becomes a single-row "big" tile (streamed in pieces by the kernel).
static void plan_tiles_of(int n_blk, int BS, const std::vector<uint64_t> &rowoff, const std::vector<int> &rows_in_blk,
                          std::vector<int2> &tiles, std::vector<int> &blk_tile_beg)
{
	blk_tile_beg.assign(n_blk + 1, 0);
	tiles.clear();
	for (int b = 0; b < n_blk; ++b) {
		blk_tile_beg[b] = (int)tiles.size();
		const uint64_t *ro = rowoff.data() + (size_t)b * (BS + 1);
		const int rows = rows_in_blk[b];
		int r = 0;
		while (r < rows) {
			if (ro[r + 1] - ro[r] > (uint64_t)RAW_CAP) { tiles.push_back(make_int2(r, (int)(1u | 0x80000000u))); ++r; continue; }
			int e = r + 1;
			while (e < rows && e - r < T_MAX && (e % COMP_K) != 0 && ro[e + 1] - ro[r] <= (uint64_t)RAW_CAP) ++e;
			tiles.push_back(make_int2(r, e - r));
			r = e;
		}
	}
	blk_tile_beg[n_blk] = (int)tiles.size();
}

static void plan_tiles(const b200_pbf_t *pb, std::vector<int2> &tiles, std::vector<int> &blk_tile_beg)
{
	plan_tiles_of(pb->n_blk, pb->BS, pb->h_rowoff, pb->rows_in_blk, tiles, blk_tile_beg);
}

// Composite maps of the row groups (compose.cu) for every sparse block: forward maps of plane 0 for the QUERY walk and
// inverse maps of the plane-1 view for the select kernel.  Query independent, so they are built once per resident PBF:
// lazily by the first split scan, or chunk by chunk while the image is still being copied (b200_pbf_load_ex).
// Which blocks are sparse is decided on the device (index.cu: p1view_kernel -> d_blk_sparse); the kernels skip the others.
static bool compose_alloc(const b200_pbf_t *pb)
{
	b200_ctx_t *c = pb->ctx;
	const int n_grp = (pb->BS + COMP_K - 1) / COMP_K;
	const size_t slots = (size_t)pb->n_blk * n_grp, vslots = (size_t)pb->n_blk * SELECT_GROUPS;
	bool ok = pool_malloc(c, (void**)&pb->d_comp_start, slots * COMP_CAP * sizeof(uint32_t)) &&
	          pool_malloc(c, (void**)&pb->d_comp_delta, slots * COMP_CAP * sizeof(int32_t)) &&
	          pool_malloc(c, (void**)&pb->d_comp_n, slots * sizeof(int) + 16) &&
	          pool_malloc(c, (void**)&pb->d_comp_dir, slots * COMP_DIR_STRIDE * sizeof(uint16_t) + 16) &&
	          pool_malloc(c, (void**)&pb->d_vcomp_start, vslots * SELECT_COMP_CAP * sizeof(uint32_t)) &&
	          pool_malloc(c, (void**)&pb->d_vcomp_delta, vslots * SELECT_COMP_CAP * sizeof(int32_t)) &&
	          pool_malloc(c, (void**)&pb->d_vcomp_n, vslots * sizeof(int) + 16) &&
	          pool_malloc(c, (void**)&pb->d_vcomp_dir, vslots * COMP_DIR_STRIDE * sizeof(uint16_t) + 16) &&
	          pool_malloc(c, (void**)&pb->d_qcol, (size_t)pb->n_blk * pb->p1_base * sizeof(int32_t) + 16) &&   // (the first p1_base pairs of every block;
	          pool_malloc(c, (void**)&pb->d_qrow, (size_t)pb->n_blk * pb->p1_base * sizeof(uint16_t) + 16) &&  //  the rest: select_extend)
	          pool_malloc(c, (void**)&pb->d_qcount, (size_t)pb->n_blk * sizeof(int) + 16);
	return ok && CU_OK(cudaMemsetAsync(pb->d_comp_n, 0, slots * sizeof(int), c->st)) && CU_OK(cudaMemsetAsync(pb->d_vcomp_n, 0, vslots * sizeof(int), c->st));
}

static ComposeParams compose_params(const b200_pbf_t *pb, int b0)
{
	ComposeParams K;
	memset(&K, 0, sizeof(K));
	K.img = pb->d_img; K.rowoff = pb->d_rowoff; K.n1 = pb->d_n1; K.rows_in_blk = pb->d_rows_in_blk; K.blk_list = nullptr; K.blk_first = b0;
	K.blk_ok = pb->d_blk_sparse; K.nrun = pb->d_nrun0;
	K.m = pb->m; K.shift = pb->shift; K.comp_start = pb->d_comp_start; K.comp_delta = pb->d_comp_delta; K.comp_n = pb->d_comp_n;
	K.n_grp = (pb->BS + COMP_K - 1) / COMP_K; K.cap = COMP_CAP; K.rle_off = 5; K.n1_plane = 0; K.inverse = 0; K.row_base = nullptr; K.n1_step = 2;
	K.comp_dir = pb->d_comp_dir; K.dir_shift = pb->dir_shift; K.dir_n = pb->dir_n;
	K.two_sided = 1; K.n_blk_res = pb->n_blk;   // inverse maps for the second half of every block that has a resident successor (pairwalk.cu)
	return K;
}

// forward composites (plane 0) of resident blocks [b0,b1): their image bytes, row index, n1 and sparse flags must be queued before
static bool compose_queue(const b200_pbf_t *pb, int b0, int b1, cudaStream_t st)
{
	if (b1 <= b0) return true;
	const bool ok = CU_OK(launch_compose(compose_params(pb, b0), b1 - b0, st));
	pb->ctx->launches += 2;
	return ok;
}

// inverse composites of the plane-1 view rows of blocks [b0,b1), then the plane-1 select over them (needs the view)
static bool select_queue(const b200_pbf_t *pb, int b0, int b1, cudaStream_t st, int *d_err)
{
	if (b1 <= b0) return true;
	b200_ctx_t *c = pb->ctx;
	ComposeParams V = compose_params(pb, b0);
	V.comp_dir = pb->d_vcomp_dir; V.nrun = nullptr; V.two_sided = 0;
	V.img = pb->d_p1img; V.rowoff = pb->d_p1_rowoff; V.n1 = pb->d_p1_n1; V.rows_in_blk = pb->d_p1_rows_in_blk;
	V.comp_start = pb->d_vcomp_start; V.comp_delta = pb->d_vcomp_delta; V.comp_n = pb->d_vcomp_n;
	V.n_grp = SELECT_GROUPS; V.cap = SELECT_COMP_CAP; V.rle_off = 9; V.n1_plane = 0; V.inverse = 1; V.row_base = pb->d_p1_vbase; V.n1_step = 1;
	SelectParams A;
	memset(&A, 0, sizeof(A));
	A.p1img = pb->d_p1img; A.p1_rowoff = pb->d_p1_rowoff; A.p1_n1 = pb->d_p1_n1; A.p1_realrow = pb->d_p1_realrow; A.p1_vbase = pb->d_p1_vbase;
	A.p1_rows_in_blk = pb->d_p1_rows_in_blk; A.img = pb->d_img; A.blkoff = pb->d_blkoff; A.blk_list = nullptr; A.blk_first = b0; A.blk_ok = pb->d_blk_sparse;
	A.m = pb->m; A.shift = pb->shift; A.cap = pb->p1_cap; A.slice0 = 0; A.n_slices = (pb->p1_base + 2047) / 2048;   // (a slice = 512 threads x 4 pairs)
	A.q_stride = pb->p1_base;
	A.vcomp_start = pb->d_vcomp_start; A.vcomp_delta = pb->d_vcomp_delta; A.vcomp_n = pb->d_vcomp_n; A.vcomp_dir = pb->d_vcomp_dir;
	A.dir_shift = pb->dir_shift; A.dir_n = pb->dir_n; A.p1_prefix = pb->d_p1_prefix;
	A.qcol = pb->d_qcol; A.qrow = pb->d_qrow; A.qcount = pb->d_qcount; A.err = d_err;
	const bool ok = CU_OK(launch_compose(V, b1 - b0, st)) && CU_OK(launch_plane1_select(A, b1 - b0, st));
	c->launches += 3;
	return ok;
}

// Blocks flagged 2 hold more pairs than the select launches of the load are sized for: the slices behind p1_base, once per resident
// PBF, sized from the real pair counts (the host has them after the load)
static bool select_extend(b200_pbf_t *pb, cudaStream_t st, int *d_err)
{
	if (pb->sel_ext_ready) return true;
	uint32_t most = 0;
	for (int b = 0; b < pb->n_blk; ++b) if (pb->blk_sparse[b] == 2 && pb->blk_ones[b] > most) most = pb->blk_ones[b];
	pb->sel_ext_ready = true;
	if (most <= (uint32_t)pb->p1_base) return true;
	b200_ctx_t *c = pb->ctx;
	long long total = 0;
	pb->ext_off.assign(pb->n_blk, 0);
	for (int b = 0; b < pb->n_blk; ++b) {
		pb->ext_off[b] = total;
		if (pb->blk_sparse[b] == 2 && pb->blk_ones[b] > (uint32_t)pb->p1_base) total += ((long long)pb->blk_ones[b] - pb->p1_base + 15) & ~15LL;
	}
	if (!pool_malloc(c, (void**)&pb->d_qcol_ext, (size_t)total * sizeof(int32_t) + 16) || !pool_malloc(c, (void**)&pb->d_qrow_ext, (size_t)total * sizeof(uint16_t) + 16) ||
	    !pool_malloc(c, (void**)&pb->d_ext_off, (size_t)pb->n_blk * sizeof(long long) + 16) ||
	    !CU_OK(cudaMemcpyAsync(pb->d_ext_off, pb->ext_off.data(), (size_t)pb->n_blk * sizeof(long long), cudaMemcpyHostToDevice, st))) { pb->sel_ext_ready = false; return false; }
	SelectParams A;
	memset(&A, 0, sizeof(A));
	A.p1img = pb->d_p1img; A.p1_rowoff = pb->d_p1_rowoff; A.p1_n1 = pb->d_p1_n1; A.p1_realrow = pb->d_p1_realrow; A.p1_vbase = pb->d_p1_vbase;
	A.p1_rows_in_blk = pb->d_p1_rows_in_blk; A.img = pb->d_img; A.blkoff = pb->d_blkoff; A.blk_list = nullptr; A.blk_first = 0; A.blk_ok = pb->d_blk_sparse;
	A.m = pb->m; A.shift = pb->shift; A.cap = pb->p1_cap; A.slice0 = pb->p1_base / 2048; A.n_slices = (int)((most + 2047) / 2048) - A.slice0;
	A.vcomp_start = pb->d_vcomp_start; A.vcomp_delta = pb->d_vcomp_delta; A.vcomp_n = pb->d_vcomp_n; A.vcomp_dir = pb->d_vcomp_dir;
	A.dir_shift = pb->dir_shift; A.dir_n = pb->dir_n; A.p1_prefix = pb->d_p1_prefix;
	A.qcol = pb->d_qcol_ext; A.qrow = pb->d_qrow_ext; A.ext_off = pb->d_ext_off; A.ext_shift = pb->p1_base; A.qcount = pb->d_qcount; A.err = d_err;
	++pb->ctx->launches;
	return CU_OK(launch_plane1_select(A, pb->n_blk, st));
}

static bool build_composites(const b200_pbf_t *pb, int *d_err)
{
	if (!compose_alloc(pb) || !compose_queue(pb, 0, pb->n_blk, pb->ctx->st) || !select_queue(pb, 0, pb->n_blk, pb->ctx->st, d_err)) return false;
	pb->comp_ready = true; pb->sel_ready = true;
	return true;
}

// Device buffers of the row index and the plane-1 view (filled by index.cu), the block table and the zeroed outputs of
// rowmeta / snapshot inversion.  The small uploads go on `up` (the copy stream when an image copy follows them).
static bool pbf_alloc_index(b200_pbf_t *pb, cudaStream_t up)
{
	b200_ctx_t *c = pb->ctx;
	const int BS = pb->BS, nb = pb->n_blk;
	const int n_grp = (BS + COMP_K - 1) / COMP_K;
	// capacity of a block's (column,row) pair list: 3 pairs per column -- a block whose plane 1 holds more ones (more than
	// one haplotype in 2700 missing / other-ALT at every site) takes the general walk
	// (r2: two tiers.  p1_base = 3 pairs per column is what the launches of the load pipeline are sized for -- nobody has seen the
	// data then, and slices without pairs cost a CTA each -- and for which the lists are allocated up front; p1_cap = 24 per column
	// (one haplotype in 340 missing / other-ALT at every site) is what a block may hold: blocks in between are flagged 2 and get
	// extension launches and extension lists sized from their real pair counts once the host has them)
	{
		static const int mult = getenv("BGT_B200_P1_CAP_MULT") ? atoi(getenv("BGT_B200_P1_CAP_MULT")) : 24;   // (tuning)
		const long long base = (((long long)pb->m * 3 > 4096 ? (long long)pb->m * 3 : 4096) + 2047) / 2048 * 2048;
		long long want = (long long)pb->m * (mult > 3 ? mult : 3);
		if (want > (1LL << 30)) want = 1LL << 30;
		if (want < base) want = base;
		pb->p1_base = (int)base;
		pb->p1_cap = (int)((want + 2047) / 2048 * 2048);
	}
	// bucket directory of the composite maps: the smallest bucket width whose directory (+ sentinel) fits COMP_DIR entries
	pb->dir_shift = 0;
	while ((((uint32_t)pb->m - 1) >> pb->dir_shift) + 2 > (uint32_t)COMP_DIR) ++pb->dir_shift;
	pb->dir_n = (int)(((((uint32_t)pb->m - 1) >> pb->dir_shift) + 2 + 7) & ~7u);
	bool ok = pool_malloc(c, (void**)&pb->d_rowoff, sizeof(uint64_t) * (size_t)nb * (BS + 1) + 8) &&
	          pool_malloc(c, (void**)&pb->d_blkoff, sizeof(uint64_t) * (nb + 1)) &&
	          pool_malloc(c, (void**)&pb->d_blkend, sizeof(uint64_t) * (nb + 1)) &&
	          pool_malloc(c, (void**)&pb->d_ix_scratch, sizeof(uint64_t) * (size_t)nb * IX_SCRATCH_LANES * (BS + 1) + 8) &&
	          pool_malloc(c, (void**)&pb->d_rows_in_blk, sizeof(int) * (nb + 1)) &&
	          pool_malloc(c, (void**)&pb->d_blk_tile_beg, sizeof(int) * (nb + 1)) &&
	          pool_malloc(c, (void**)&pb->d_blk_tile_end, sizeof(int) * (nb + 1)) &&
	          pool_malloc(c, (void**)&pb->d_grp_tile_beg, sizeof(int) * (size_t)nb * (n_grp + 1) + 16) &&
	          pool_malloc(c, (void**)&pb->d_tiles, sizeof(int2) * ((size_t)nb * BS + 1)) &&
	          pool_malloc(c, (void**)&pb->d_n1, sizeof(uint32_t) * (size_t)nb * BS * 2 + 8) &&
	          pool_malloc(c, (void**)&pb->d_nrun0, sizeof(uint32_t) * (size_t)nb * BS + 8) &&
	          pool_malloc(c, (void**)&pb->d_rank0, sizeof(int32_t) * (size_t)nb * 2 * (size_t)pb->m + 8) &&
	          pool_malloc(c, (void**)&pb->d_p1img, (size_t)nb * P1_SLOT_BYTES + 64) &&
	          pool_malloc(c, (void**)&pb->d_p1_rowoff, sizeof(uint64_t) * (size_t)nb * (SELECT_MAX_ROWS + 1) + 8) &&
	          pool_malloc(c, (void**)&pb->d_p1_n1, sizeof(uint32_t) * (size_t)nb * SELECT_MAX_ROWS + 8) &&
	          pool_malloc(c, (void**)&pb->d_p1_prefix, sizeof(uint32_t) * (size_t)nb * (SELECT_MAX_ROWS + 1) + 8) &&
	          pool_malloc(c, (void**)&pb->d_p1_realrow, sizeof(uint16_t) * (size_t)nb * SELECT_MAX_ROWS + 8) &&
	          pool_malloc(c, (void**)&pb->d_p1_rows_in_blk, sizeof(int) * (nb + 1)) &&
	          pool_malloc(c, (void**)&pb->d_p1_vbase, sizeof(long long) * (nb + 1)) &&
	          pool_malloc(c, (void**)&pb->d_blk_sparse, (size_t)nb + 16) && pool_malloc(c, (void**)&pb->d_blk_ones, sizeof(uint32_t) * (size_t)nb + 16);
	if (!ok) return false;
	return CU_OK(cudaMemsetAsync(c->d_acc + 4, 0, 2 * sizeof(unsigned long long), up)) &&
	       CU_OK(cudaMemsetAsync(c->d_err, 0, sizeof(int), up)) &&
	       CU_OK(cudaMemcpyAsync(pb->d_blkoff, pb->h_blkoff.data(), sizeof(uint64_t) * nb, cudaMemcpyHostToDevice, up)) &&
	       CU_OK(cudaMemcpyAsync(pb->d_blkend, pb->h_blkend.data(), sizeof(uint64_t) * nb, cudaMemcpyHostToDevice, up)) &&
	       CU_OK(cudaMemcpyAsync(pb->d_rows_in_blk, pb->rows_in_blk.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, up)) &&
	       CU_OK(cudaMemsetAsync(pb->d_n1, 0, sizeof(uint32_t) * (size_t)nb * BS * 2, c->st)) &&
	       CU_OK(cudaMemsetAsync(pb->d_rank0, 0, sizeof(int32_t) * (size_t)nb * 2 * (size_t)pb->m, c->st));
}

static IndexParams index_params(const b200_pbf_t *pb, int b0)
{
	IndexParams X;
	memset(&X, 0, sizeof(X));
	X.img = pb->d_img; X.blkoff = pb->d_blkoff; X.blkend = pb->d_blkend; X.rows_in_blk = pb->d_rows_in_blk;
	X.m = pb->m; X.shift = pb->shift; X.blk_first = b0; X.rowoff = pb->d_rowoff; X.scratch = pb->d_ix_scratch; X.tiles = pb->d_tiles;
	X.blk_tile_beg = pb->d_blk_tile_beg; X.blk_tile_end = pb->d_blk_tile_end; X.grp_tile_beg = pb->d_grp_tile_beg; X.err = pb->ctx->d_err; X.fallbacks = (int*)(pb->ctx->d_acc + 5);
	return X;
}

// blocks [b0,b1) whose row offsets are known (pbf_index_kernel): tiles and per-row n1
static bool queue_tiles_rowmeta(b200_pbf_t *pb, int b0, int b1, cudaStream_t st)
{
	b200_ctx_t *c = pb->ctx;
	const int BS = pb->BS;
	c->launches += 2;
	return CU_OK(launch_plan_tiles(index_params(pb, b0), b1 - b0, st)) &&
	       CU_OK(launch_rowmeta(pb->d_img, pb->d_rowoff + (size_t)b0 * (BS + 1), b1 - b0, pb->shift, 0, pb->d_rows_in_blk + b0, (uint32_t)pb->m,
	                            pb->d_n1 + (size_t)b0 * BS * 2, pb->d_nrun0 + (size_t)b0 * BS, c->d_acc + 4, st));
}

// blocks [b0,b1) whose snapshots are in place: start ranks (pbwt.c:343) and the plane-1 view
static bool queue_ranks_view(b200_pbf_t *pb, int b0, int b1, cudaStream_t st)
{
	b200_ctx_t *c = pb->ctx;
	P1ViewParams V;
	memset(&V, 0, sizeof(V));
	V.img = pb->d_img; V.rowoff = pb->d_rowoff; V.n1 = pb->d_n1; V.rows_in_blk = pb->d_rows_in_blk;
	V.m = pb->m; V.shift = pb->shift; V.blk_first = b0; V.p1_cap = pb->p1_cap; V.p1_base = pb->p1_base; V.blk_ones = pb->d_blk_ones;
	V.p1img = pb->d_p1img; V.p1_rowoff = pb->d_p1_rowoff; V.p1_n1 = pb->d_p1_n1; V.p1_realrow = pb->d_p1_realrow; V.p1_prefix = pb->d_p1_prefix;
	V.p1_rows_in_blk = pb->d_p1_rows_in_blk; V.p1_vbase = pb->d_p1_vbase; V.blk_sparse = pb->d_blk_sparse;
	c->launches += 2;
	return CU_OK(launch_invert_snapshots(pb->d_img, pb->d_blkoff + b0, b1 - b0, pb->m, pb->d_rank0 + (size_t)b0 * 2 * (size_t)pb->m, c->d_err, st)) &&
	       CU_OK(launch_p1view(V, b1 - b0, st));
}

// End of a load: fetch the per-block flags and the error state, wait for everything queued.
static bool pbf_finish_load(b200_pbf_t *pb)
{
	b200_ctx_t *c = pb->ctx;
	unsigned long long bad = 0;
	int err = 0;
	pb->blk_sparse.assign(pb->n_blk ? pb->n_blk : 1, 0);
	bool ok = CU_OK(cudaMemcpyAsync(&bad, c->d_acc + 4, sizeof(bad), cudaMemcpyDeviceToHost, c->st)) &&
	          CU_OK(cudaMemcpyAsync(&err, c->d_err, sizeof(err), cudaMemcpyDeviceToHost, c->st));
	pb->blk_ones.assign(pb->n_blk ? pb->n_blk : 1, 0);
	if (ok && pb->n_blk) ok = CU_OK(cudaMemcpyAsync(pb->blk_sparse.data(), pb->d_blk_sparse, (size_t)pb->n_blk, cudaMemcpyDeviceToHost, c->st)) &&
	                          CU_OK(cudaMemcpyAsync(pb->blk_ones.data(), pb->d_blk_ones, sizeof(uint32_t) * (size_t)pb->n_blk, cudaMemcpyDeviceToHost, c->st));
	ok = ok && CU_OK(cudaStreamSynchronize(c->st));
	if (!ok) return false;
	pb->bad_rows = (int64_t)bad;
	if (getenv("BGT_B200_TRACE")) {
		unsigned long long fb = 0;
		cudaMemcpy(&fb, c->d_acc + 5, sizeof(fb), cudaMemcpyDeviceToHost);
		fprintf(stderr, "[b200 trace]   row index: %llu block(s) re-chased by a single lane\n", fb & 0xffffffffull);
	}
	if (err & 128) { set_err("the records of one checkpoint block exceed 4 GiB; not supported"); return false; }
	if (err & 2) { set_err("corrupt PBF: record tags/lengths inside a checkpoint block do not parse"); return false; }
	if (err & 1) { set_err("corrupt PBF: an 'S' snapshot holds a column index >= m"); return false; }
	if (err & 8) { set_err("internal: TMA copy never completed"); return false; }
	pb->p1_ready = true;
	return true;
}

// host copy of the device-built row offsets, for the few host-side questions that need them (b200_pbf_row_bytes)
static const std::vector<uint64_t> *host_rowoff(const b200_pbf_t *cpb)
{
	b200_pbf_t *pb = const_cast<b200_pbf_t*>(cpb);
	if (pb->h_rowoff.empty() && pb->n_blk) {
		pb->h_rowoff.assign((size_t)pb->n_blk * (pb->BS + 1), 0);
		cudaSetDevice(pb->ctx->dev);
		if (!CU_OK(cudaMemcpyAsync(pb->h_rowoff.data(), pb->d_rowoff, pb->h_rowoff.size() * 8, cudaMemcpyDeviceToHost, pb->ctx->st)) ||
		    !CU_OK(cudaStreamSynchronize(pb->ctx->st))) { pb->h_rowoff.clear(); return nullptr; }
	}
	return &pb->h_rowoff;
}



// the file indexes checkpoint blocks only (pbwt.c:297); walk the length prefixes of the rows inside
static bool walk_block(const uint8_t *f, size_t flen, uint64_t off, int m, int g, int rows, uint64_t *roff)
{
	if (off >= flen || f[off] != 'S') return false;
	uint64_t pos = off + 1 + (uint64_t)g * 4 * (uint64_t)m;
	for (int r = 0; r < rows; ++r) {
		if (pos >= flen || f[pos] != 'B') return false;
		roff[r] = pos++;
		for (int p = 0; p < g; ++p) {
			int32_t l;
			if (pos + 4 > flen) return false;
			memcpy(&l, f + pos, 4);
			if (l < 0 || pos + 4 + (uint64_t)l > flen) return false;
			pos += 4 + (uint64_t)l;
		}
	}
	roff[rows] = pos;
	return true;
}

// Host half of making rows [row_beg,row_end) resident: header / index parse (pbwt.c:231-258), block range,
// per-block row offsets (a few host threads).  Offsets are relative to pb->file_off0.
static b200_pbf_t *pbf_index_prepare(const uint8_t *f, size_t flen, int64_t row_beg, int64_t row_end)
{
	if (!f) { set_err("null PBF image"); return nullptr; }
	if (flen < 16 + 13 + 8 || memcmp(f, "PBF\1", 4) != 0) { set_err("not a PBF file (bad magic)"); return nullptr; } // pbwt.c:231-235
	int32_t hdr[3];
	memcpy(hdr, f + 4, 12);
	uint64_t ioff;
	memcpy(&ioff, f + flen - 8, 8);
	if (ioff > flen - 13 - 8 || f[ioff] != 'I') { set_err_code(B200_E_CORRUPT, "PBF has no index record; it cannot be made resident"); return nullptr; } // pbwt.c:247-258 (no wrap-around: ioff is untrusted)
	int64_t n; int32_t n_idx;
	memcpy(&n, f + ioff + 1, 8);
	memcpy(&n_idx, f + ioff + 9, 4);
	const int m = hdr[0], g = hdr[1], shift = hdr[2];
	if (g != 2) { set_err("PBF has %d bit planes; the BGT genotype path uses exactly 2 (import.c:68)", g); return nullptr; }
	if (m <= 0 || m >= (1 << 30) || shift < 0 || shift > 24 || n < 0 || n_idx < 0 || (uint64_t)n_idx > (flen - ioff - 13) / 8 || n > ((int64_t)n_idx << shift)) { set_err("PBF header out of range (m=%d shift=%d n=%lld)", m, shift, (long long)n); return nullptr; }
	const int BS = 1 << shift;
	if ((int64_t)n_idx != (n + BS - 1) / BS) { set_err("PBF index has %d entries for %lld rows", n_idx, (long long)n); return nullptr; }
	if (row_end < 0 || row_end > n) row_end = n;
	if (row_beg < 0) row_beg = 0;
	if (row_beg > row_end) row_beg = row_end;
	std::vector<uint64_t> idx(n_idx);
	memcpy(idx.data(), f + ioff + 13, 8ull * n_idx);
	for (int32_t i = 0; i < n_idx; ++i)
		if (idx[i] < 16 || idx[i] >= ioff || (i && idx[i] <= idx[i - 1])) { set_err("corrupt PBF: block index entry %d out of order", i); return nullptr; }

	b200_pbf_t *pb = new b200_pbf_t();
	pb->m = m; pb->g = g; pb->shift = shift; pb->BS = BS; pb->n = n; pb->n_blk_file = n_idx;
	pb->blk0 = (int)(row_beg >> shift);
	const int blk1 = row_end > row_beg ? (int)((row_end + BS - 1) >> shift) : pb->blk0;
	pb->n_blk = blk1 - pb->blk0;
	const int nb = pb->n_blk;
	const uint64_t byte_beg = nb ? idx[pb->blk0] : ioff, byte_end = (nb && blk1 < n_idx) ? idx[blk1] : ioff;
	pb->file_off0 = byte_beg & ~(uint64_t)15;
	if (pb->blk0 == 0 && blk1 == n_idx) { pb->file_off0 = 0; pb->file_size = flen; }
	const uint64_t copy_end = pb->file_size ? flen : byte_end;
	pb->img_bytes = (size_t)(copy_end - pb->file_off0);
	pb->h_idx.swap(idx);
	pb->ioff = ioff;
	// block table: everything the device indexer needs comes from the header and the index record
	pb->rows_in_blk.resize(nb);
	pb->h_blkoff.resize(nb);
	pb->h_blkend.resize(nb);
	for (int b = 0; b < nb; ++b) {
		const int64_t r0 = (int64_t)(pb->blk0 + b) << shift;
		pb->rows_in_blk[b] = (int)(n - r0 < BS ? n - r0 : BS);
		pb->h_blkoff[b] = pb->h_idx[pb->blk0 + b] - pb->file_off0;
		pb->h_blkend[b] = (pb->blk0 + b + 1 < n_idx ? pb->h_idx[pb->blk0 + b + 1] : ioff) - pb->file_off0;
		if (pb->h_blkoff[b] + 1 + (uint64_t)g * 4 * (uint64_t)m > pb->h_blkend[b]) { set_err("corrupt PBF: checkpoint block %d is shorter than its 'S' record", pb->blk0 + b); delete pb; return nullptr; }
	}
	return pb;
}

// Host mirror of index.cu's row index, for b200_pbf_plan only (planning / validation without a device): walk the
// length prefixes of every resident block (the file indexes only block starts, pbwt.c:297).
static bool pbf_index_walk(b200_pbf_t *pb, const uint8_t *f)
{
	const int nb = pb->n_blk, BS = pb->BS, shift = pb->shift, m = pb->m, g = pb->g;
	const int64_t n = pb->n;
	const std::vector<uint64_t> &idx = pb->h_idx;
	const uint64_t ioff = pb->ioff;
	pb->h_rowoff.assign((size_t)nb * (BS + 1), 0);
	(void)shift; (void)n;
	bool ok = true;
	{
		int hw = (int)std::thread::hardware_concurrency();
		hw = hw < 8 ? 8 : (hw > 32 ? 32 : hw);
		const int nt = nb < 2 ? 1 : (nb < hw ? nb : hw);
		std::vector<int> bad(nt, 0);
		std::vector<std::thread> th;
		for (int t = 0; t < nt; ++t)
			th.emplace_back([&, t]() {
				for (int b = t; b < nb; b += nt) {
					uint64_t *ro = pb->h_rowoff.data() + (size_t)b * (BS + 1);
					if (!walk_block(f, (size_t)ioff + 1, idx[pb->blk0 + b], m, g, pb->rows_in_blk[b], ro)) { bad[t] = 1; return; }
					for (int r = 0; r <= pb->rows_in_blk[b]; ++r) ro[r] -= pb->file_off0;
				}
			});
		for (auto &x : th) x.join();
		for (int t = 0; t < nt; ++t) if (bad[t]) ok = false;
	}
	if (!ok) { set_err("corrupt PBF: record tags/lengths inside a checkpoint block do not parse"); return false; }
	return true;
}

static b200_pbf_t *pbf_index_host(const uint8_t *f, size_t flen, int64_t row_beg, int64_t row_end)
{
	b200_pbf_t *pb = pbf_index_prepare(f, flen, row_beg, row_end);
	if (pb && !pbf_index_walk(pb, f)) { delete pb; return nullptr; }
	return pb;
}


// Fused load + count scan (b200_pbf_load_scan): what the chunk pipeline of the load queues behind every chunk's composite maps
struct FusedScan {
	const b200_query_t *q;
	b200_scan_out_t *out;       // host outputs (counts, pass)
	int64_t row_lo, row_hi;     // rows to produce (clipped to the file by the caller)
	int32_t *d_counts; uint8_t *d_pass;
	bool queued = false;        // pair walk + finalize + D2H were queued for every resident block
};
static PairParams pair_params(b200_ctx_t *c, const b200_pbf_t *pb, const b200_query_t *q, int64_t row_lo, int64_t row_hi);

static b200_pbf_t *pbf_load_impl(b200_ctx_t *c, const uint8_t *f, size_t flen, int64_t row_beg, int64_t row_end, unsigned flags, FusedScan *fs)
{
	if (!c || !f) { set_err("b200_pbf_load: null argument"); return nullptr; }
	cudaSetDevice(c->dev);
	const bool trace = getenv("BGT_B200_TRACE") != nullptr;
	const double t0 = now_ms();
	b200_pbf_t *pb = pbf_index_prepare(f, flen, row_beg, row_end);
	if (!pb) return nullptr;
	pb->ctx = c;
	pb->prepare_split = (flags & B200_LOAD_PREPARE_COUNT_SCAN) != 0;
	// The host only parses the header and the block index (pbwt.c:231-258) and queues work; the rows inside the blocks
	// are indexed on the device (index.cu).  The image goes out in up to LOAD_CHUNKS chunks of whole checkpoint blocks on the copy
	// stream; behind every chunk's event: the length-prefix chase of its blocks (st_idx, latency bound, so it runs beside
	// the previous chunk's kernels), then tiles, row meta, start ranks, the plane-1 view and -- if asked for -- the
	// composite maps on st.  One synchronisation at the end.
	const int nb = pb->n_blk;
	bool ok = pool_malloc(c, (void**)&pb->d_img, pb->img_bytes + 64);
	ok = ok && CU_OK(cudaStreamSynchronize(c->st));          // buffers from the pool may still be in use by queued work of this context
	ok = ok && pbf_alloc_index(pb, c->st_copy);
	const bool eager = pb->prepare_split;
	const bool copy_only = getenv("BGT_B200_COPYONLY") != nullptr;   // diagnostics: time the bare H2D copy
	if (eager) ok = ok && compose_alloc(pb);
	// fused scan: per-site raw counters of the rows to produce, zeroed before the first pair walk is queued
	PairParams FK;
	FinalizeSplit fsp;
	memset(&FK, 0, sizeof(FK)); memset(&fsp, 0, sizeof(fsp));
	int64_t f_lo = 0, f_hi = 0;
	if (fs && ok) {
		f_lo = fs->row_lo; f_hi = fs->row_hi < pb->n ? fs->row_hi : pb->n;
		const size_t nr = (size_t)(f_hi > f_lo ? f_hi - f_lo : 0);
		ok = eager && nb > 0 && nr > 0 && c->cnt_raw.reserve(nr * 3 * sizeof(int32_t)) && c->counts.reserve(nr * 6 * sizeof(int32_t)) && c->pass.reserve(nr);
		if (ok) {
			fs->d_counts = (int32_t*)c->counts.p; fs->d_pass = (uint8_t*)c->pass.p;
			ok = CU_OK(cudaMemsetAsync(c->cnt_raw.p, 0, nr * 3 * sizeof(int32_t), c->st)) && CU_OK(cudaMemsetAsync(c->d_acc, 0, 4 * sizeof(unsigned long long), c->st)) &&
			     CU_OK(cudaMemsetAsync(c->d_err_scan, 0, sizeof(int), c->st)) &&
			     CU_OK(cudaEventRecord(c->ev_comp[LOAD_CHUNKS], c->st)) && CU_OK(cudaStreamWaitEvent(c->st_walk, c->ev_comp[LOAD_CHUNKS], 0));
			FK = pair_params(c, pb, fs->q, f_lo, f_hi);
			FK.blk_ok = pb->d_blk_sparse;
			fsp.blk_split = pb->d_blk_sparse; fsp.n1 = pb->d_n1; fsp.blk_row0 = (long long)pb->blk0 << pb->shift; fsp.shift = pb->shift;
		} else if (pb) { set_err("b200_pbf_load_scan: nothing to scan"); }
	}
	int fused_done_blk = 0;     // resident blocks [0, fused_done_blk) have their pair walk queued
	auto fused_queue = [&](int hi, int k) { // pair walk of resident blocks [fused_done_blk, hi), their finalize, their results home
		if (!fs || hi <= fused_done_blk) return true;
		const int lo = fused_done_blk;
		fused_done_blk = hi;
		PairParams K = FK;
		K.blk_list = nullptr; K.blk_first = lo;
		bool good = CU_OK(launch_pairwalk(K, 4, pb->p1_base, hi - lo, c->st_walk));   // (blocks flagged 2 are redone by the ordinary scan)
		++c->launches;
		const long long blk_row0 = (long long)pb->blk0 << pb->shift;
		long long r0 = blk_row0 + ((long long)lo << pb->shift), r1 = blk_row0 + ((long long)hi << pb->shift);
		if (r0 < f_lo) r0 = f_lo;
		if (r1 > f_hi) r1 = f_hi;
		if (good && r1 > r0) {
			const size_t off = (size_t)(r0 - f_lo), cnt = (size_t)(r1 - r0);
			FinalizeSplit sp = fsp;
			sp.row_lo = r0;
			good = CU_OK(launch_finalize((const int32_t*)c->cnt_raw.p + off * 3, (long long)cnt, 1, fs->q->d_gsize, fs->q->d_prog, fs->q->has_flt, fs->d_counts + off * 6, fs->d_pass + off,
			                             c->d_acc, sp, c->st_walk)) &&
			       CU_OK(cudaEventRecord(c->ev_fin[k], c->st_walk)) && CU_OK(cudaStreamWaitEvent(c->st_d2h, c->ev_fin[k], 0));
			++c->launches;
			if (good && fs->out->counts) good = CU_OK(cudaMemcpyAsync(fs->out->counts + off * 6, fs->d_counts + off * 6, cnt * 6 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st_d2h));
			if (good && fs->out->pass) good = CU_OK(cudaMemcpyAsync(fs->out->pass + off, fs->d_pass + off, cnt, cudaMemcpyDeviceToHost, c->st_d2h));
		}
		return good;
	};
	// the zeroing memsets on st must be done before kernels on the index streams write n1 / rank0
	ok = ok && CU_OK(cudaEventRecord(c->ev_zero, c->st));
	for (int i = 0; ok && i < N_IDX_STREAMS; ++i) ok = CU_OK(cudaStreamWaitEvent(c->st_idx[i], c->ev_zero, 0));
	ok = ok && CU_OK(cudaEventRecord(c->ev[4], c->st_copy));
	const int n_chunks = nb >= 2 * LOAD_CHUNKS ? LOAD_CHUNKS : (nb >= 16 ? 8 : (nb > 0 ? 1 : 0));
	size_t done = 0;
	cudaEvent_t tev[LOAD_CHUNKS][4];
	bool comp2_used = false;
	if (n_chunks == 0) ok = ok && CU_OK(cudaMemsetAsync(pb->d_img, 0, 64, c->st_copy));
	// chunk bounds: even, except that a long image starts with a ramp of short chunks (1, 2, 3, 4, 6, 8 parts of 128) so that
	// the first index chain -- and with it the composite maps, which everything else queues behind -- starts early and
	// every later chain is hidden behind the composite maps of the chunk before; and ends with the mirrored ramp, so that
	// what is left to do behind the last chunk's copy (its maps, its pair walk, its results) is short
	int cb[LOAD_CHUNKS + 1];
	{
		static const int ramp[6] = {1, 2, 3, 4, 6, 8};
		// (short chunks at the END as well shorten what is left to do behind the last copy -- which pays when the COPY is the limit
		// (several GPUs sharing the host link) and costs when the SMs are: measured on one GPU 11.3 ms with a full mirrored ramp of
		// 24 chunks, 10.5 ms with three short chunks at the end, 9.9 ms without.  Opt-in: BGT_B200_TAIL_RAMP=1.)
		const int head = (n_chunks == LOAD_CHUNKS && nb >= 64) ? 6 : 0, tail = (head && getenv("BGT_B200_TAIL_RAMP")) ? 3 : 0;
		int h = 0, t = nb;
		cb[0] = 0; cb[n_chunks] = nb;
		for (int k = 0; k < head; ++k) { const int sz = nb * ramp[k] / 128; h += sz > 0 ? sz : 1; cb[k + 1] = h; }
		for (int k = 0; k < tail; ++k) { const int sz = nb * ramp[k] / 128; t -= sz > 0 ? sz : 1; cb[n_chunks - 1 - k] = t; }
		const int mid = n_chunks - head - tail;
		for (int k = 1; k < mid; ++k) cb[head + k] = h + (int)((long long)(t - h) * k / mid);
	}
	for (int k = 0; ok && k < n_chunks; ++k) {
		const int b0 = cb[k], b1 = cb[k + 1];
		const size_t end = b1 < nb ? (size_t)(pb->h_idx[pb->blk0 + b1] - pb->file_off0) : pb->img_bytes;
		ok = CU_OK(cudaMemcpyAsync(pb->d_img + done, f + pb->file_off0 + done, end - done, cudaMemcpyHostToDevice, c->st_copy));
		if (ok && k == n_chunks - 1) ok = CU_OK(cudaMemsetAsync(pb->d_img + pb->img_bytes, 0, 64, c->st_copy));
		ok = ok && CU_OK(cudaEventRecord(c->ev_chunk[k], c->st_copy));
		if (trace) { cudaEventCreate(&tev[k][3]); cudaEventRecord(tev[k][3], c->st_copy); }
		done = end;
		if (copy_only) continue;
		// per chunk, on one of the index streams: chase -> tiles, n1, start ranks, plane-1 view (small latency-bound grids that
		// overlap the previous chunks' composite maps, which are the only thing left on st)
		cudaStream_t sx = c->st_idx[k % N_IDX_STREAMS];
		ok = ok && CU_OK(cudaStreamWaitEvent(sx, c->ev_chunk[k], 0)) && CU_OK(launch_index(index_params(pb, b0), b1 - b0, sx));
		++c->launches;
		ok = ok && queue_tiles_rowmeta(pb, b0, b1, sx) && queue_ranks_view(pb, b0, b1, sx);
		static const bool one_comp_stream = getenv("BGT_B200_ONE_COMP_STREAM") != nullptr;   // (A/B measurements)
		cudaStream_t sc = (eager && (k & 1) && !one_comp_stream) ? c->st_comp2 : c->st;
		if (sc != c->st && !comp2_used) { comp2_used = true; ok = ok && CU_OK(cudaEventRecord(c->ev_comp2, c->st)) && CU_OK(cudaStreamWaitEvent(sc, c->ev_comp2, 0)); }   // (behind whatever was queued on st before this load)
		ok = ok && CU_OK(cudaEventRecord(c->ev_idx[k], sx)) && CU_OK(cudaStreamWaitEvent(sc, c->ev_idx[k], 0));
		if (eager) { // forward composites on the main stream and its twin, alternating; the plane-1 side (inverse composites + select: small grids) stays on the index stream
			ok = ok && compose_queue(pb, b0, b1, sc) && select_queue(pb, b0, b1, sx, c->d_err) &&
			     CU_OK(cudaEventRecord(c->ev_sel[k], sx));
			if (fs && ok) { // the pair walk of a block reads the start ranks of the block behind it (two-sided maps): the chunk's last
			                // block waits for the next chunk
				ok = CU_OK(cudaEventRecord(c->ev_comp[k], sc)) && CU_OK(cudaStreamWaitEvent(c->st_walk, c->ev_comp[k], 0)) &&
				     CU_OK(cudaStreamWaitEvent(c->st_walk, c->ev_sel[k], 0));
				// (a chunk's pair walk is fewer CTAs than the device holds, each alive for hundreds of microseconds with 110 KB of shared
				// memory beside the composite maps of the next chunks.  Queueing it behind every second chunk -- fewer, fuller launches --
				// measured 8.45 against 8.64 ms per step (two boxes, interleaved runs); every 4th: no gain, every 8th: 9.1 ms, the
				// last walk behind the last copy gets too long.  BGT_B200_PAIR_EVERY for A/B runs.)
				static const int pair_every = getenv("BGT_B200_PAIR_EVERY") ? atoi(getenv("BGT_B200_PAIR_EVERY")) : 2;
				if (ok && ((k + 1) % (pair_every > 0 ? pair_every : 1) == 0 || k == n_chunks - 1)) ok = fused_queue(k == n_chunks - 1 ? b1 : b1 - 1, k);
			}
		}
		if (trace) {
			for (int j = 0; j < 3; ++j) cudaEventCreate(&tev[k][j]);
			cudaEventRecord(tev[k][1], sx); cudaEventRecord(tev[k][2], sc);
		}
	}
	ok = ok && CU_OK(cudaEventRecord(c->ev[5], c->st_copy));
	if (comp2_used) ok = ok && CU_OK(cudaEventRecord(c->ev_comp2, c->st_comp2)) && CU_OK(cudaStreamWaitEvent(c->st, c->ev_comp2, 0));   // st covers its twin again
	if (eager && !copy_only) for (int k = 0; ok && k < n_chunks; ++k) ok = CU_OK(cudaStreamWaitEvent(c->st, c->ev_sel[k], 0));
	if (fs && ok && !copy_only) {
		ok = fused_queue(nb, LOAD_CHUNKS); fs->queued = ok && fused_done_blk == nb;
		// the final synchronisation of the load (st) covers the walks as well
		ok = ok && CU_OK(cudaEventRecord(c->ev_fin[LOAD_CHUNKS], c->st_walk)) && CU_OK(cudaStreamWaitEvent(c->st, c->ev_fin[LOAD_CHUNKS], 0));
	}
	const double t1 = now_ms();
	ok = ok && pbf_finish_load(pb);
	if (ok && eager) { pb->comp_ready = true; pb->sel_ready = true; }
	if (trace) {
		fprintf(stderr, "[b200 trace] load: queued in %.2f ms, device done after %.2f ms\n", t1 - t0, now_ms() - t0);
		for (int k = 0; k < n_chunks && !copy_only; ++k) {
			float a = 0, b = 0, d = 0;
			cudaEventElapsedTime(&a, c->ev[4], tev[k][3]); cudaEventElapsedTime(&b, c->ev[4], tev[k][1]); cudaEventElapsedTime(&d, c->ev[4], tev[k][2]);
			fprintf(stderr, "[b200 trace]   chunk %2d: copied at %.2f ms, indexed at %.2f, composites at %.2f\n", k, a, b, d);
			for (int j = 0; j < 4; ++j) cudaEventDestroy(tev[k][j]);
		}
	}
	if (!ok) { cudaStreamSynchronize(c->st_copy); for (int i = 0; i < N_IDX_STREAMS; ++i) cudaStreamSynchronize(c->st_idx[i]); cudaStreamSynchronize(c->st_walk); cudaStreamSynchronize(c->st_comp2); cudaStreamSynchronize(c->st); cudaStreamSynchronize(c->st_d2h); pbf_free_device(pb); delete pb; return nullptr; }
	float ms = 0;
	if (cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]) == cudaSuccess) c->last_ms[2] = ms;
	return pb;
}

extern "C" b200_pbf_t *b200_pbf_load_ex(b200_ctx_t *c, const uint8_t *f, size_t flen, int64_t row_beg, int64_t row_end, unsigned flags)
{
	return pbf_load_impl(c, f, flen, row_beg, row_end, flags, nullptr);
}

extern "C" b200_pbf_t *b200_pbf_load(b200_ctx_t *c, const uint8_t *f, size_t flen, int64_t row_beg, int64_t row_end)
{
	return b200_pbf_load_ex(c, f, flen, row_beg, row_end, 0);
}

extern "C" int b200_pbf_plan(const uint8_t *f, size_t flen, int64_t row_beg, int64_t row_end, int64_t info[8])
{
	b200_pbf_t *pb = pbf_index_host(f, flen, row_beg, row_end);
	if (!pb) return -1;
	std::vector<int2> tiles;
	std::vector<int> btb;
	plan_tiles(pb, tiles, btb);
	int64_t n_big = 0, max_tile = 0, max_row = 0;
	for (int b = 0; b < pb->n_blk; ++b) {
		const uint64_t *ro = pb->h_rowoff.data() + (size_t)b * (pb->BS + 1);
		for (int t = btb[b]; t < btb[b + 1]; ++t) {
			const int r = tiles[t].x, nr = tiles[t].y & 0x7fffffff;
			const int64_t bytes = (int64_t)(ro[r + nr] - ro[r]);
			if (tiles[t].y < 0) ++n_big; else if (bytes > max_tile) max_tile = bytes;
		}
		for (int r = 0; r < pb->rows_in_blk[b]; ++r) if ((int64_t)(ro[r + 1] - ro[r]) > max_row) max_row = (int64_t)(ro[r + 1] - ro[r]);
	}
	if (info) {
		info[0] = pb->m; info[1] = pb->shift; info[2] = pb->n; info[3] = pb->n_blk; info[4] = (int64_t)tiles.size();
		info[5] = n_big; info[6] = max_tile; info[7] = max_row;
	}
	delete pb;
	return 0;
}

extern "C" int b200_flt_eval_host(const char *flt, int n_groups, const int32_t *counts, int64_t n_rows, uint8_t *pass)
{
	if (n_groups < 1 || n_groups > B200_MAX_GROUPS) { set_err("n_groups out of range"); return -1; }
	flt_prog_t prog;
	const int err = flt_compile(flt, n_groups, &prog);
	if (err) { set_err("filter expression does not parse (kexpr error mask 0x%x)", err); return err; }
	const int stride = 3 + 3 * n_groups;
	for (int64_t k = 0; k < n_rows; ++k) pass[k] = (uint8_t)flt_eval(&prog, counts + k * stride);
	return 0;
}

extern "C" b200_pbf_t *b200_pbf_open(b200_ctx_t *c, const char *fn, int64_t row_beg, int64_t row_end)
{
	if (!fn) { set_err("b200_pbf_open: null file name"); return nullptr; }
	const int fd = open(fn, O_RDONLY);
	if (fd < 0) { set_err("cannot open '%s'", fn); return nullptr; }
	struct stat sb;
	if (fstat(fd, &sb) != 0 || sb.st_size < 16) { close(fd); set_err("'%s' is not a PBF file", fn); return nullptr; }
	void *mp = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
	close(fd);
	if (mp == MAP_FAILED) { set_err("mmap of '%s' failed", fn); return nullptr; }
	b200_pbf_t *pb = b200_pbf_load(c, (const uint8_t*)mp, (size_t)sb.st_size, row_beg, row_end);
	munmap(mp, (size_t)sb.st_size);
	return pb;
}

extern "C" int b200_pbf_m(const b200_pbf_t *pb) { return pb ? pb->m : -1; }
extern "C" int b200_pbf_g(const b200_pbf_t *pb) { return pb ? pb->g : -1; }
extern "C" int b200_pbf_shift(const b200_pbf_t *pb) { return pb ? pb->shift : -1; }
extern "C" int64_t b200_pbf_n(const b200_pbf_t *pb) { return pb ? pb->n : -1; }
extern "C" int64_t b200_pbf_row_beg(const b200_pbf_t *pb) { return pb ? (int64_t)pb->blk0 << pb->shift : -1; }
extern "C" int64_t b200_pbf_row_end(const b200_pbf_t *pb)
{
	if (!pb) return -1;
	const int64_t e = (int64_t)(pb->blk0 + pb->n_blk) << pb->shift;
	return e < pb->n ? e : pb->n;
}
extern "C" int64_t b200_pbf_bad_rows(const b200_pbf_t *pb) { return pb ? pb->bad_rows : -1; }
extern "C" int b200_pbf_split_blocks(const b200_pbf_t *pb)
{
	if (!pb) return -1;
	int k = 0;
	for (int b = 0; b < pb->n_blk && b < (int)pb->blk_sparse.size(); ++b) k += pb->blk_sparse[b] != 0;
	return k;
}
extern "C" size_t b200_pbf_image_size(const b200_pbf_t *pb) { return pb ? pb->file_size : 0; }

extern "C" int64_t b200_pbf_row_bytes(const b200_pbf_t *pb, int64_t row_beg, int64_t row_end, int with_snapshots)
{
	if (!pb) return -1;
	if (row_beg < b200_pbf_row_beg(pb) || row_end > b200_pbf_row_end(pb) || row_beg > row_end) { set_err("row range not resident"); return -1; }
	int64_t bytes = 0;
	const int BS = pb->BS;
	const std::vector<uint64_t> *hro = host_rowoff(pb);
	if (!hro) return -1;
	for (int64_t k = row_beg; k < row_end;) {
		const int b = (int)(k >> pb->shift) - pb->blk0, r = (int)(k & (BS - 1));
		const int64_t in_blk = pb->rows_in_blk[b] - r;
		const int64_t take = row_end - k < in_blk ? row_end - k : in_blk;
		const uint64_t *ro = hro->data() + (size_t)b * (BS + 1);
		bytes += (int64_t)(ro[r + take] - ro[r]);
		if (with_snapshots && r == 0) bytes += 1 + (int64_t)pb->g * 4 * pb->m;
		k += take;
	}
	return bytes;
}

// file byte range [*byte_beg, *byte_end) of the checkpoint blocks that hold rows [row_beg,row_end): what a region shard uploads
// (b200_pbf_load_ex reads the header, this range and the index record at the tail of the image, nothing else)
extern "C" int b200_pbf_block_bytes(const b200_pbf_t *pb, int64_t row_beg, int64_t row_end, uint64_t *byte_beg, uint64_t *byte_end, uint64_t *index_beg)
{
	if (!pb || !byte_beg || !byte_end) { set_err("b200_pbf_block_bytes: null argument"); return -1; }
	if (row_end < 0 || row_end > pb->n) row_end = pb->n;
	if (row_beg < 0 || row_beg > row_end) { set_err("b200_pbf_block_bytes: bad row range"); return -1; }
	const int64_t b0 = row_beg >> pb->shift, b1 = row_end > row_beg ? ((row_end + pb->BS - 1) >> pb->shift) : b0;
	if (b1 > (int64_t)pb->h_idx.size()) { set_err("b200_pbf_block_bytes: rows beyond the block index"); return -1; }
	*byte_beg = b0 < (int64_t)pb->h_idx.size() ? pb->h_idx[(size_t)b0] : pb->ioff;
	*byte_end = b1 < (int64_t)pb->h_idx.size() ? pb->h_idx[(size_t)b1] : pb->ioff;
	if (index_beg) *index_beg = pb->ioff;
	return 0;
}

// part of the file image of a fully resident PBF (generated cohorts): bytes [off, off + n_bytes) -> dst
extern "C" int b200_pbf_image_download_range(const b200_pbf_t *pb, uint8_t *dst, uint64_t off, size_t n_bytes)
{
	if (!pb || !dst) return -1;
	if (!pb->file_size || off > pb->file_size || n_bytes > pb->file_size - off) { set_err("no complete file image resident, or range outside it"); return -1; }
	cudaSetDevice(pb->ctx->dev);
	if (!CU_OK(cudaMemcpyAsync(dst, pb->d_img + off, n_bytes, cudaMemcpyDeviceToHost, pb->ctx->st))) return -1;
	return CU_OK(cudaStreamSynchronize(pb->ctx->st)) ? 0 : -1;
}

extern "C" int b200_pbf_image_download(const b200_pbf_t *pb, uint8_t *dst, size_t n_bytes)
{
	if (!pb || !dst) return -1;
	if (!pb->file_size || n_bytes < pb->file_size) { set_err("no complete file image resident, or buffer too small"); return -1; }
	cudaSetDevice(pb->ctx->dev);
	if (!CU_OK(cudaMemcpyAsync(dst, pb->d_img, pb->file_size, cudaMemcpyDeviceToHost, pb->ctx->st))) return -1;
	return CU_OK(cudaStreamSynchronize(pb->ctx->st)) ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------ query

extern "C" void b200_query_destroy(b200_query_t *q)
{
	if (!q) return;
	cudaSetDevice(q->ctx->dev);
	cudaStreamSynchronize(q->ctx->st);
	if (q->d_track) cudaFree(q->d_track);
	if (q->d_tgrp) cudaFree(q->d_tgrp);
	if (q->d_gsize) cudaFree(q->d_gsize);
	if (q->d_prog) cudaFree(q->d_prog);
	delete q;
}

extern "C" b200_query_t *b200_query_create(b200_ctx_t *c, const b200_pbf_t *pb, int n_out, const int32_t *out_samples,
                                           const uint32_t *group, int n_groups, const char *flt, int *flt_err)
{
	if (flt_err) *flt_err = 0;
	if (!c || !pb) { set_err("b200_query_create: null argument"); return nullptr; }
	return b200_query_create_m(c, pb->m, n_out, out_samples, group, n_groups, flt, flt_err);
}

// the same for a PBF that is not resident yet (b200_pbf_load_scan): m = its number of columns (b200_pbf_peek)
extern "C" b200_query_t *b200_query_create_m(b200_ctx_t *c, int m, int n_out, const int32_t *out_samples,
                                             const uint32_t *group, int n_groups, const char *flt, int *flt_err)
{
	if (flt_err) *flt_err = 0;
	if (!c || m <= 0) { set_err("b200_query_create: null argument"); return nullptr; }
	cudaSetDevice(c->dev);
	const int n_samples = m / 2;
	if (n_groups < 1 || n_groups > B200_MAX_GROUPS) { set_err("n_groups=%d out of range 1..%d", n_groups, B200_MAX_GROUPS); return nullptr; }
	if (out_samples == nullptr) n_out = n_samples;
	if (n_out < 0 || n_out > n_samples) { set_err("n_out=%d out of range", n_out); return nullptr; }
	b200_query_t *q = new b200_query_t();
	q->ctx = c; q->m = m; q->n_out = n_out; q->n_track = 2 * n_out; q->G = n_groups;
	q->words = (q->n_track + 31) / 32;
	// pbwt.c:377: asking for >= m columns means "decode everything"; out[] is ascending, so that is the identity
	q->full = (q->n_track >= m);
	std::vector<int32_t> track(q->n_track ? q->n_track : 1);
	std::vector<uint8_t> tgrp(q->n_track ? q->n_track : 1);
	std::vector<int32_t> gsize(B200_MAX_GROUPS, 0);
	for (int i = 0; i < n_out; ++i) {
		const int s = out_samples ? out_samples[i] : i;
		if (s < 0 || s >= n_samples || (i && out_samples && s <= out_samples[i - 1])) { set_err("out_samples must be ascending sample indices in 0..%d", n_samples - 1); delete q; return nullptr; }
		const uint32_t gr = group ? group[i] : 1u;
		if (gr < 1 || gr > (uint32_t)n_groups) { set_err("group[%d]=%u out of range 1..%d", i, gr, n_groups); delete q; return nullptr; }
		track[2 * i] = 2 * s; track[2 * i + 1] = 2 * s + 1;      // bgt.c:240-242
		tgrp[2 * i] = tgrp[2 * i + 1] = (uint8_t)(gr - 1);       // bgt.c:743-744
		gsize[gr - 1] += 2;
	}
	const int perr = flt_compile(flt, n_groups, &q->prog);
	if (perr) { if (flt_err) *flt_err = perr; set_err_code(B200_E_FILTER_SYNTAX, "filter expression does not parse (kexpr error mask 0x%x)", perr); delete q; return nullptr; }
	q->has_flt = q->prog.n > 0;
	bool ok = CU_OK(cudaMalloc(&q->d_tgrp, tgrp.size() + 16)) && CU_OK(cudaMalloc(&q->d_gsize, sizeof(int32_t) * B200_MAX_GROUPS)) &&
	          CU_OK(cudaMalloc(&q->d_prog, sizeof(flt_prog_t)));
	if (ok && !q->full) ok = CU_OK(cudaMalloc(&q->d_track, sizeof(int32_t) * track.size()));
	ok = ok && CU_OK(cudaMemcpyAsync(q->d_tgrp, tgrp.data(), tgrp.size(), cudaMemcpyHostToDevice, c->st)) &&
	     CU_OK(cudaMemcpyAsync(q->d_gsize, gsize.data(), sizeof(int32_t) * B200_MAX_GROUPS, cudaMemcpyHostToDevice, c->st)) &&
	     CU_OK(cudaMemcpyAsync(q->d_prog, &q->prog, sizeof(flt_prog_t), cudaMemcpyHostToDevice, c->st));
	if (ok && !q->full) ok = CU_OK(cudaMemcpyAsync(q->d_track, track.data(), sizeof(int32_t) * track.size(), cudaMemcpyHostToDevice, c->st));
	ok = ok && CU_OK(cudaStreamSynchronize(c->st));
	if (!ok) { b200_query_destroy(q); return nullptr; }
	return q;
}

extern "C" b200_query_t *b200_query_create_cols(b200_ctx_t *c, const b200_pbf_t *pb, int n_cols, const int32_t *cols)
{
	if (!c || !pb) { set_err("b200_query_create_cols: null argument"); return nullptr; }
	cudaSetDevice(c->dev);
	if (n_cols <= 0 || n_cols >= pb->m || cols == nullptr) { n_cols = pb->m; cols = nullptr; } // pbwt.c:377
	b200_query_t *q = new b200_query_t();
	q->ctx = c; q->m = pb->m; q->n_out = n_cols / 2; q->n_track = n_cols; q->G = 1; q->words = (n_cols + 31) / 32;
	q->full = (cols == nullptr);
	memset(&q->prog, 0, sizeof(q->prog));
	std::vector<int32_t> gsize(B200_MAX_GROUPS, 0);
	gsize[0] = n_cols;
	for (int i = 0; cols && i < n_cols; ++i)
		if (cols[i] < 0 || cols[i] >= pb->m) { set_err("column %d out of range 0..%d", cols[i], pb->m - 1); delete q; return nullptr; }
	bool ok = CU_OK(cudaMalloc(&q->d_tgrp, (size_t)n_cols + 16)) && CU_OK(cudaMalloc(&q->d_gsize, sizeof(int32_t) * B200_MAX_GROUPS)) &&
	          CU_OK(cudaMalloc(&q->d_prog, sizeof(flt_prog_t)));
	if (ok && cols) ok = CU_OK(cudaMalloc(&q->d_track, sizeof(int32_t) * (size_t)n_cols));
	ok = ok && CU_OK(cudaMemsetAsync(q->d_tgrp, 0, (size_t)n_cols + 16, c->st)) &&
	     CU_OK(cudaMemcpyAsync(q->d_gsize, gsize.data(), sizeof(int32_t) * B200_MAX_GROUPS, cudaMemcpyHostToDevice, c->st)) &&
	     CU_OK(cudaMemcpyAsync(q->d_prog, &q->prog, sizeof(flt_prog_t), cudaMemcpyHostToDevice, c->st));
	if (ok && cols) ok = CU_OK(cudaMemcpyAsync(q->d_track, cols, sizeof(int32_t) * (size_t)n_cols, cudaMemcpyHostToDevice, c->st));
	ok = ok && CU_OK(cudaStreamSynchronize(c->st));
	if (!ok) { b200_query_destroy(q); return nullptr; }
	return q;
}

extern "C" int b200_query_n_track(const b200_query_t *q) { return q ? q->n_track : -1; }
extern "C" int b200_query_filter_needs_host(const b200_query_t *q) { return q && q->has_flt && q->prog.needs_host ? 1 : 0; }
extern "C" int b200_query_hap_words(const b200_query_t *q) { return q ? q->words : -1; }
extern "C" int b200_query_counts_stride(const b200_query_t *q) { return q ? 3 + 3 * q->G : -1; }

// ------------------------------------------------------------------------------------------------ scan

// the pair walk (pairwalk.cu) over the resident PBF's own pair lists and composite maps
static PairParams pair_params(b200_ctx_t *c, const b200_pbf_t *pb, const b200_query_t *q, int64_t row_lo, int64_t row_hi)
{
	PairParams K;
	memset(&K, 0, sizeof(K));
	K.img = pb->d_img; K.rowoff = pb->d_rowoff; K.n1 = pb->d_n1; K.rank0 = pb->d_rank0;
	K.qcol = pb->d_qcol; K.qrow = pb->d_qrow; K.qcount = pb->d_qcount; K.q_stride = pb->p1_base; K.tgrp = q->d_tgrp;
	K.comp_start = pb->d_comp_start; K.comp_delta = pb->d_comp_delta; K.comp_n = pb->d_comp_n; K.comp_dir = pb->d_comp_dir;
	K.dir_shift = pb->dir_shift; K.dir_n = pb->dir_n; K.cnt_raw = (int32_t*)c->cnt_raw.p;
	K.m = pb->m; K.G = q->G; K.shift = pb->shift; K.blk_row0 = (long long)pb->blk0 << pb->shift; K.row_lo = row_lo; K.row_hi = row_hi; K.err = c->d_err_scan;
	K.rows_in_blk = pb->d_rows_in_blk; K.two_sided = 1; K.n_blk_res = pb->n_blk;
	return K;
}

static int pick_cols_per_thread(const b200_ctx_t *c, int n_track, int n_blk)
{
	// enough CTAs to fill the chip twice over (2 resident CTAs per SM), else fewer columns per thread
	for (int C = 8; C > 1; C >>= 1) {
		const long long ctas = (long long)((n_track + WALK_NT * C - 1) / (WALK_NT * C)) * n_blk;
		if (ctas >= 4LL * c->sm_count && n_track >= WALK_NT * C) return C;
	}
	return 1;
}

// device times of the last scan: [0] all phases between the first and the last walk-type kernel, [1] whole scan,
// [4] plane1_select_kernel, [5] pbwt_marginal_kernel; the rank-walk kernel alone is [0] - [4] - [5]
static void read_scan_timings(b200_ctx_t *c)
{
	float ms;
	c->last_ms[4] = c->last_ms[5] = 0;
	if (cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) c->last_ms[0] = ms;
	if (cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]) == cudaSuccess) c->last_ms[1] = ms;
	if (c->split_used && cudaEventElapsedTime(&ms, c->ev[8], c->ev[9]) == cudaSuccess) c->last_ms[4] = ms;
	if (c->marginal_used && cudaEventElapsedTime(&ms, c->ev[10], c->ev[11]) == cudaSuccess) c->last_ms[5] = ms;
}

extern "C" int64_t b200_scan(b200_ctx_t *c, const b200_pbf_t *pb, const b200_query_t *q, int64_t row_beg, int64_t n_rows,
                             unsigned flags, b200_scan_out_t *out)
{
	if (!c || !pb || !q || !out) { set_err("b200_scan: null argument"); return -1; }
	if (pb->ctx != c || q->ctx != c) { set_err("b200_scan: handles belong to another context"); return -1; }
	if (q->m != pb->m) { set_err("b200_scan: query was built for m=%d, PBF has m=%d", q->m, pb->m); return -1; }
	cudaSetDevice(c->dev);
	if (row_beg < 0 || n_rows < 0) { set_err("b200_scan: negative row range"); return -1; }
	if (row_beg + n_rows > pb->n) n_rows = pb->n - row_beg > 0 ? pb->n - row_beg : 0;
	out->totals[0] = out->totals[1] = out->totals[2] = out->totals[3] = 0;
	if (n_rows == 0) return 0;
	if (row_beg < b200_pbf_row_beg(pb) || row_beg + n_rows > b200_pbf_row_end(pb)) { set_err("b200_scan: rows [%lld,%lld) are not resident", (long long)row_beg, (long long)(row_beg + n_rows)); return -1; }
	const bool dev_out = flags & B200_SCAN_DEVICE_OUT;
	const bool want_bits = (flags & B200_SCAN_HAP_BITS) && out->hap_bits[0] && out->hap_bits[1];
	const bool want_bytes = (flags & B200_SCAN_HAP_BYTES) && out->hap_bytes[0] && out->hap_bytes[1];
	const bool emit = want_bits || want_bytes;
	const bool want_counts = (flags & B200_SCAN_COUNTS) && out->counts;
	const bool host_flt = q->has_flt && q->prog.needs_host;
	if (host_flt && dev_out) { set_err_code(B200_E_FILTER_NEEDS_HOST, "filters using ** are evaluated with the host libm; not available with B200_SCAN_DEVICE_OUT"); return -1; }
	const int G = q->G, stride = 3 + 3 * G, words = q->words, n_track = q->n_track;
	const size_t nr = (size_t)n_rows;

	// ---- device outputs: the caller's buffers (DEVICE_OUT) or context scratch
	int32_t *d_counts = nullptr; uint8_t *d_pass = nullptr; uint32_t *d_bits[2] = {nullptr, nullptr}; uint8_t *d_bytes[2] = {nullptr, nullptr};
	if (!c->cnt_raw.reserve(nr * G * 3 * sizeof(int32_t))) return -1;
	if (dev_out) {
		d_counts = out->counts; d_pass = out->pass;
		if (want_bits) { d_bits[0] = out->hap_bits[0]; d_bits[1] = out->hap_bits[1]; }
		if (want_bytes) { d_bytes[0] = out->hap_bytes[0]; d_bytes[1] = out->hap_bytes[1]; }
	} else {
		if (want_counts || host_flt) { if (!c->counts.reserve(nr * stride * sizeof(int32_t))) return -1; d_counts = (int32_t*)c->counts.p; }
		if (out->pass) { if (!c->pass.reserve(nr)) return -1; d_pass = (uint8_t*)c->pass.p; }
	}
	if (emit && !d_bits[0]) {
		for (int p = 0; p < 2; ++p) { if (!c->hapbits[p].reserve(nr * words * sizeof(uint32_t))) return -1; d_bits[p] = (uint32_t*)c->hapbits[p].p; }
	}
	if (want_bytes && !d_bytes[0]) {
		for (int p = 0; p < 2; ++p) { if (!c->hapbytes[p].reserve(nr * (size_t)n_track)) return -1; d_bytes[p] = (uint8_t*)c->hapbytes[p].p; }
	}

	// ---- launch
	const int b_first = (int)(row_beg >> pb->shift) - pb->blk0;
	const int b_last = (int)((row_beg + n_rows - 1) >> pb->shift) - pb->blk0;
	WalkParams P;
	memset(&P, 0, sizeof(P));
	P.img = pb->d_img; P.rowoff = pb->d_rowoff; P.n1 = pb->d_n1; P.tiles = pb->d_tiles; P.blk_tile_beg = pb->d_blk_tile_beg; P.blk_tile_end = pb->d_blk_tile_end;
	P.rank0 = pb->d_rank0; P.track = q->full ? nullptr : q->d_track; P.tgrp = q->d_tgrp;
	P.cnt_raw = (int32_t*)c->cnt_raw.p; P.hap[0] = d_bits[0]; P.hap[1] = d_bits[1];
	P.m = pb->m; P.n_track = n_track; P.G = G; P.words = words; P.shift = pb->shift;
	P.blk_first = b_first; P.blk_row0 = (long long)pb->blk0 << pb->shift; P.row_lo = row_beg; P.row_hi = row_beg + n_rows; P.err = c->d_err_scan;
	const int forced = (int)((flags >> 8) & 15u);

	// Split scan (all columns, one group, counts only): #ALT of a site is the number of ones of its plane-0 row, known
	// from the RLE alone; only the columns that carry a missing / other-ALT code somewhere in the block (the set W,
	// from the block's plane-1 rows) have to be walked to resolve the joint codes.  Blocks whose W would be large
	// (pb->blk_sparse == 0) and every other kind of query take the general walk over all tracked columns.
	// (several groups: the per-group marginals come from marginal.cu -- one bit vector per group in shared memory)
	const bool split = q->full && !emit && pb->p1_ready && !(flags & B200_SCAN_NO_SPLIT) && n_track > 0 &&
	                   (G == 1 || (G <= 8 && marginal_smem_bytes(pb->m) <= 200 * 1024));
	// Which block takes which path was decided on the device while loading (d_blk_sparse; the host holds a copy): every kernel
	// is launched over the whole block range of the scan and the CTAs of the other path's blocks exit -- no per-scan lists to
	// upload (an asynchronous upload from a host vector must not outlive it, and B200_SCAN_DEVICE_OUT scans return at once).
	int n_gen = 0, n_split = 0;
	for (int b = b_first; b <= b_last; ++b) { if (split && pb->blk_sparse[b]) ++n_split; else ++n_gen; }
	const int n_range = b_last - b_first + 1;
	const uint8_t *d_split_flags = split ? pb->d_blk_sparse : nullptr;
	FinalizeSplit sp;
	memset(&sp, 0, sizeof(sp));

	bool ok = CU_OK(cudaEventRecord(c->ev[2], c->st)) &&
	          (c->fused_pairs_done || CU_OK(cudaMemsetAsync(c->cnt_raw.p, 0, nr * G * 3 * sizeof(int32_t), c->st))) &&
	          (c->keep_totals || CU_OK(cudaMemsetAsync(c->d_acc, 0, 4 * sizeof(unsigned long long), c->st))) &&
	          (c->fused_pairs_done || c->keep_totals || CU_OK(cudaMemsetAsync(c->d_err_scan, 0, sizeof(int), c->st))) &&
	          CU_OK(cudaEventRecord(c->ev[0], c->st));
	if (ok && n_track > 0 && n_gen > 0) {
		P.blk_list = nullptr; P.blk_first = b_first; P.blk_skip = d_split_flags;
		const int Cg = (forced == 1 || forced == 2 || forced == 4 || forced == 8) ? forced : pick_cols_per_thread(c, n_track, n_gen);
		ok = CU_OK(launch_walk(P, Cg, emit ? WALK_MODE_EMIT : WALK_MODE_COUNT, (n_track + WALK_NT * Cg - 1) / (WALK_NT * Cg), n_range, c->st));
		c->launches += (n_range + 32767) / 32768;
	}
	c->split_used = n_split > 0; c->marginal_used = n_split > 0 && G > 1;
	if (ok && n_split > 0) {
		const int cap = pb->p1_cap;
		if (!pb->comp_ready && !(flags & B200_SCAN_NO_COMPOSE)) { // composite maps of the row groups: once per resident PBF
			if (!build_composites(pb, c->d_err_scan)) return -1;
		}
		// phase 1 (plane1.cu): the (column, row) pairs that carry a plane-1 bit, per block, in row order.  They do not depend
		// on the query: found once per resident PBF next to the composite maps (select_queue); only the testing path without
		// composites runs the select per scan.
		const bool use_comp = pb->comp_ready && pb->sel_ready && !(flags & B200_SCAN_NO_COMPOSE);
		const int32_t *qcol = pb->d_qcol; const uint16_t *qrow = pb->d_qrow; const int *qcount = pb->d_qcount;
		if (!use_comp) {
			if (!c->qcol.reserve((size_t)pb->n_blk * cap * sizeof(int32_t)) || !c->qrow.reserve((size_t)pb->n_blk * cap * sizeof(uint16_t)) ||
			    !c->qcount.reserve((size_t)pb->n_blk * sizeof(int))) return -1;
			SelectParams A;
			memset(&A, 0, sizeof(A));
			A.p1img = pb->d_p1img; A.p1_rowoff = pb->d_p1_rowoff; A.p1_n1 = pb->d_p1_n1; A.p1_realrow = pb->d_p1_realrow; A.p1_vbase = pb->d_p1_vbase;
			A.p1_prefix = pb->d_p1_prefix; A.dir_shift = pb->dir_shift; A.dir_n = pb->dir_n;
			A.p1_rows_in_blk = pb->d_p1_rows_in_blk; A.img = pb->d_img; A.blkoff = pb->d_blkoff; A.blk_list = nullptr; A.blk_first = b_first; A.blk_ok = d_split_flags;
			A.m = pb->m; A.shift = pb->shift; A.cap = cap;
			A.qcol = (int32_t*)c->qcol.p; A.qrow = (uint16_t*)c->qrow.p; A.qcount = (int*)c->qcount.p; A.err = c->d_err_scan;
			ok = CU_OK(cudaEventRecord(c->ev[8], c->st)) && CU_OK(launch_plane1_select(A, n_range, c->st)) && CU_OK(cudaEventRecord(c->ev[9], c->st));
			++c->launches;
			qcol = (const int32_t*)c->qcol.p; qrow = (const uint16_t*)c->qrow.p; qcount = (const int*)c->qcount.p;
		} else c->split_used = false;   // (no select time in this scan)
		// phase 2: walk those haplotypes through plane 0 up to their row and add code 3 (other-ALT) or 2 (missing) there
		WalkParams B = P;
		if (use_comp) {
			B.comp_start = pb->d_comp_start; B.comp_delta = pb->d_comp_delta; B.comp_n = pb->d_comp_n; B.grp_tile_beg = pb->d_grp_tile_beg;
			B.comp_dir = pb->d_comp_dir; B.dir_shift = pb->dir_shift; B.dir_n = pb->dir_n;
		}
		B.track = qcol; B.qrow = qrow; B.track_stride = cap;
		B.n_track_blk = qcount; B.n_track = cap; B.blk_list = nullptr; B.blk_first = b_first; B.blk_skip = nullptr; B.blk_ok = d_split_flags;
		const int Cb = (forced == 1 || forced == 2 || forced == 4 || forced == 8) ? forced : (use_comp ? 2 : 8);
		if (use_comp && c->fused_pairs_done) {
			// (b200_pbf_load_scan: the pair walk of every block was queued behind the block's composite maps while loading)
		} else if (use_comp) {   // (without composite maps -- testing -- the QUERY mode of the walk kernel goes row by row)
			PairParams K = pair_params(c, pb, q, row_beg, row_beg + n_rows);
			K.blk_list = nullptr; K.blk_first = b_first; K.blk_ok = d_split_flags;
			static const bool prof = getenv("BGT_B200_PROF") != nullptr;
			static unsigned long long *d_prof = nullptr;
			if (prof) {
				if (!d_prof) cudaMalloc(&d_prof, 8 * sizeof(unsigned long long));
				cudaMemsetAsync(d_prof, 0, 8 * sizeof(unsigned long long), c->st);
				K.prof = d_prof;
			}
			static const int env_c = getenv("BGT_B200_PAIR_C") ? atoi(getenv("BGT_B200_PAIR_C")) : 0;   // tuning
			const int Cp = (forced == 1 || forced == 2 || forced == 4) ? forced : (env_c == 1 || env_c == 2 || env_c == 4 ? env_c : 4);
			// grid from the real pair counts (the host has them since the load); the slices behind p1_base -- blocks flagged 2 -- in a
			// second launch, after the select has been extended to them
			uint32_t most = 0; bool ext = false;
			for (int b = b_first; b <= b_last; ++b) if (pb->blk_sparse[b]) { if (pb->blk_ones[b] > most) most = pb->blk_ones[b]; ext = ext || pb->blk_sparse[b] == 2; }
			const int base_pairs = most < (uint32_t)pb->p1_base ? (int)most : pb->p1_base;
			if (base_pairs > 0) ok = ok && CU_OK(launch_pairwalk(K, Cp, base_pairs, n_range, c->st));
			if (ext && most > (uint32_t)pb->p1_base) {
				ok = ok && select_extend(const_cast<b200_pbf_t*>(pb), c->st, c->d_err_scan);
				K.slice0 = pb->p1_base / (PAIR_SLICE_THREADS * Cp);
				K.qcol = pb->d_qcol_ext; K.qrow = pb->d_qrow_ext; K.ext_off = pb->d_ext_off; K.ext_shift = pb->p1_base;
				ok = ok && CU_OK(launch_pairwalk(K, Cp, (int)(most - (uint32_t)pb->p1_base), n_range, c->st));
				++c->launches;
			}
			if (prof) {
				unsigned long long h[8];
				cudaMemcpyAsync(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost, c->st);
				cudaStreamSynchronize(c->st);
				fprintf(stderr, "[b200 prof] pair walk: %llu CTAs, mean cycles phase A %.0f (%.1f groups), phase B %.0f (%.1f rows per warp)\n", h[2], (double)h[0] / (h[2] ? h[2] : 1),
				        (double)h[3] / (h[2] ? h[2] : 1), (double)h[1] / (h[2] ? h[2] : 1), (double)h[4] / (h[2] ? h[2] * 16.0 : 1));
			}
		} else
			ok = ok && CU_OK(launch_walk(B, Cb, WALK_MODE_QUERY, (cap + WALK_NT * Cb - 1) / (WALK_NT * Cb), n_range, c->st));
		++c->launches;
		if (G > 1) { // per-group plane-0 marginals for the first G-1 groups (the last one is the remainder)
			if (!c->n0g.reserve(nr * (size_t)(G - 1) * sizeof(int32_t))) return -1;
			MarginalParams M;
			memset(&M, 0, sizeof(M));
			M.n_seg = 1; M.seg_slots = 1; M.seg_slot_step = 1;
			M.img = pb->d_img; M.rowoff = pb->d_rowoff; M.n1 = pb->d_n1; M.blkoff = pb->d_blkoff; M.rows_in_blk = pb->d_rows_in_blk; M.nrun = pb->d_nrun0;
			M.blk_list = nullptr; M.blk_first = b_first; M.blk_ok = d_split_flags; M.tgrp = q->d_tgrp; M.n0g = (int32_t*)c->n0g.p; M.m = pb->m; M.shift = pb->shift; M.n_vec = G - 1;
			M.blk_row0 = P.blk_row0; M.row_lo = row_beg; M.row_hi = row_beg + n_rows;
			const int n_grp = (pb->BS + COMP_K - 1) / COMP_K;
			const size_t seg_bytes = marginal_seg_words(pb->m) * sizeof(uint32_t);
			const bool segments = use_comp && !(flags & B200_SCAN_NO_SEGMENTS);
			const bool pieces = segments && !(flags & B200_SCAN_NO_PIECES) && pb->BS % (8 * COMP_K) == 0 && n_grp < 65536;
			ok = ok && CU_OK(cudaEventRecord(c->ev[10], c->st));
			if (pieces) {
				// margpiece.cu: a vector in front of every 32-row group (and behind the last full one) with prefix counts, then two CTAs per
				// group that carry the MAP of up to 16 rows instead of the vector; blocks it cannot hold are redone by the row loop from
				// every 8th vector.  Batches of blocks keep the vectors within 4 GB.
				const size_t rec_bytes = marginal_rec_words64(pb->m) * 16;
				const size_t per_blk = (size_t)(G - 1) * (size_t)(n_grp + 1) * rec_bytes;
				int batch = (int)std::max<size_t>(1, ((size_t)4 << 30) / per_blk);
				if (batch > n_range) batch = n_range;
				const size_t n_flag = ((size_t)batch * (size_t)(G - 1) + 63) & ~(size_t)63;
				const int retry_cap = 1 << 16;
				if (!c->vseg.reserve(per_blk * (size_t)batch + 256) || !c->seg_ok.reserve(2 * n_flag + 64 + (size_t)retry_cap * sizeof(uint2))) return -1;
				M.comp_start = pb->d_comp_start; M.comp_delta = pb->d_comp_delta; M.comp_n = pb->d_comp_n; M.two_sided = 1; M.n_blk_res = pb->n_blk;
				M.n_grp = n_grp; M.vseg = nullptr; M.vrec = (uint4*)c->vseg.p;
				M.seg_ok = (uint8_t*)c->seg_ok.p; M.blk_fail = M.seg_ok + n_flag; M.retry_n = (int*)(M.blk_fail + n_flag); M.retry = (uint2*)(M.blk_fail + n_flag + 64); M.retry_cap = retry_cap;
				M.seg_slots = n_grp + 1;
				const size_t vec_words = (size_t)(G - 1) * (size_t)(n_grp + 1) * marginal_rec_words64(pb->m);   // uint4 per block
				for (int b0 = 0; b0 < n_range && ok; b0 += batch) {
					const int nb = std::min(batch, n_range - b0);
					M.blk_first = b_first + b0;
					M.seg_groups = 1; M.n_seg = n_grp; M.seg_slot_step = 1; M.blk_only = nullptr;
					ok = CU_OK(cudaMemsetAsync(M.seg_ok, 1, n_flag, c->st)) && CU_OK(cudaMemsetAsync(M.blk_fail, 0, n_flag + 64, c->st));
					// The last block of the scan usually has nothing behind it to walk backward from (no resident successor, or the scan ends
					// inside it): its vectors are one chain of up to 256 steps, twice as long as everybody else's two half chains.  It runs on a
					// second stream beside the other blocks' vectors and chains.
					const bool side = nb >= 2 && b0 + nb == n_range;
					MarginalParams M2 = M;
					const int nb1 = side ? nb - 1 : nb;
					if (side) {
						M2.blk_first = M.blk_first + nb1; M2.vrec = M.vrec + (size_t)nb1 * vec_words;
						M2.seg_ok = M.seg_ok + (size_t)nb1 * (size_t)(G - 1); M2.blk_fail = M.blk_fail + (size_t)nb1 * (size_t)(G - 1);
						M2.retry_n = M.retry_n + 1; M2.retry = M.retry + retry_cap / 2; M2.retry_cap = retry_cap / 2; M.retry_cap = retry_cap / 2;
						ok = ok && CU_OK(cudaEventRecord(c->ev_fin[0], c->st)) && CU_OK(cudaStreamWaitEvent(c->st_walk, c->ev_fin[0], 0)) &&
						     CU_OK(launch_marginal_dense_seed(M2, 1, c->st_walk)) && CU_OK(cudaEventRecord(c->ev_fin[1], c->st_walk));
					}
					ok = ok && CU_OK(launch_marginal_dense_seed(M, nb1, c->st)) && CU_OK(launch_marginal_pieces(M, nb1, c->st));
					if (side) ok = ok && CU_OK(cudaStreamWaitEvent(c->st, c->ev_fin[1], 0)) && CU_OK(launch_marginal_pieces(M2, 1, c->st));
					M.seg_groups = 8; M.n_seg = n_grp / 8; M.seg_slot_step = 8; M.blk_only = M.blk_fail; M.retry_cap = retry_cap;
					ok = ok && CU_OK(launch_marginal_rows(M, nb, c->st));     // (flags and vectors of both parts lie back to back)
					c->launches += side ? 7 : 4;
				}
			} else {
				if (segments) { // segments of 8+ row groups, as many as 512 MB of segment vectors allow
					int seg_groups = 8;
					const size_t per_seg = (size_t)n_range * (size_t)(G - 1) * seg_bytes;
					while (seg_groups < n_grp && per_seg * (size_t)((n_grp + seg_groups - 1) / seg_groups) > ((size_t)512 << 20)) seg_groups *= 2;
					const int n_seg = (n_grp + seg_groups - 1) / seg_groups;
					if (n_seg > 1) {
						if (!c->vseg.reserve(per_seg * (size_t)n_seg) || !c->seg_ok.reserve((size_t)n_range * (size_t)(G - 1) + 16)) return -1;
						M.comp_start = pb->d_comp_start; M.comp_delta = pb->d_comp_delta; M.comp_n = pb->d_comp_n; M.two_sided = 1; M.n_blk_res = pb->n_blk;
						M.n_grp = n_grp; M.seg_groups = seg_groups; M.n_seg = n_seg; M.vseg = (uint32_t*)c->vseg.p; M.seg_ok = (uint8_t*)c->seg_ok.p;
						M.seg_slots = n_seg; M.seg_slot_step = 1;
						++c->launches;
					}
				}
				ok = ok && CU_OK(launch_marginal(M, n_range, c->st));
				++c->launches;
			}
			ok = ok && CU_OK(cudaEventRecord(c->ev[11], c->st));
			sp.n0g = (const int32_t*)c->n0g.p; sp.n_vec = G - 1;
		}
		sp.blk_split = d_split_flags; sp.n1 = pb->d_n1; sp.row_lo = row_beg; sp.blk_row0 = P.blk_row0; sp.shift = pb->shift;
	}
	ok = ok && CU_OK(cudaEventRecord(c->ev[1], c->st));
	ok = ok && CU_OK(launch_finalize(P.cnt_raw, n_rows, G, q->d_gsize, q->d_prog, q->has_flt && !host_flt, d_counts, d_pass, c->d_acc, sp, c->st));
	++c->launches;
	if (ok && want_bytes) {
		for (int p = 0; p < 2 && ok; ++p) { ok = CU_OK(launch_unpack_bits(d_bits[p], n_rows, words, n_track, d_bytes[p], c->st)); ++c->launches; }
	}
	ok = ok && CU_OK(cudaEventRecord(c->ev[3], c->st));
	if (!ok) return -1;
	if (dev_out) return n_rows; // asynchronous: totals are not available in this mode

	// ---- results to the host
	unsigned long long tot[4] = {0, 0, 0, 0};
	int err = 0;
	ok = CU_OK(cudaEventRecord(c->ev[6], c->st));
	if (ok && (want_counts || host_flt)) {
		int32_t *dst = want_counts ? out->counts : nullptr;
		std::vector<int32_t> tmp;
		if (!dst) { tmp.resize(nr * stride); dst = tmp.data(); }
		ok = CU_OK(cudaMemcpyAsync(dst, d_counts, nr * stride * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
		if (ok && host_flt) { // `**` goes through the host libm (kexpr.c:150), on the device-computed counts
			ok = CU_OK(cudaStreamSynchronize(c->st));
			unsigned long long np = 0;
			for (size_t k = 0; ok && k < nr; ++k) { const int pass = flt_eval(&q->prog, dst + k * stride); if (out->pass) out->pass[k] = (uint8_t)pass; np += pass; }
			tot[3] = np;
		}
	}
	if (ok && out->pass && !host_flt) ok = CU_OK(cudaMemcpyAsync(out->pass, d_pass, nr, cudaMemcpyDeviceToHost, c->st));
	for (int p = 0; p < 2 && ok; ++p) {
		if (want_bits) ok = CU_OK(cudaMemcpyAsync(out->hap_bits[p], d_bits[p], nr * words * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st));
		if (ok && want_bytes) ok = CU_OK(cudaMemcpyAsync(out->hap_bytes[p], d_bytes[p], nr * (size_t)n_track, cudaMemcpyDeviceToHost, c->st));
	}
	unsigned long long dtot[4];
	ok = ok && CU_OK(cudaMemcpyAsync(dtot, c->d_acc, sizeof(dtot), cudaMemcpyDeviceToHost, c->st)) &&
	     CU_OK(cudaMemcpyAsync(&err, c->d_err_scan, sizeof(err), cudaMemcpyDeviceToHost, c->st)) &&
	     CU_OK(cudaEventRecord(c->ev[7], c->st)) && CU_OK(cudaStreamSynchronize(c->st));
	if (!ok) return -1;
	if (err) { set_err("device error flags 0x%x during scan", err); return -1; }
	out->totals[0] = (int64_t)dtot[0]; out->totals[1] = (int64_t)dtot[1]; out->totals[2] = (int64_t)dtot[2];
	out->totals[3] = host_flt ? (int64_t)tot[3] : (int64_t)dtot[3];
	for (int i = 0; i < 4; ++i) c->last_totals[i] = out->totals[i];
	read_scan_timings(c);
	float ms;
	if (cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]) == cudaSuccess) c->last_ms[3] = ms;
	return n_rows;
}

// Batched regions: what the reference does with one pbf_seek + pbf_read loop per region (pbwt.c:349-372; bgt.c:333-345 behind
// `-r` / `-B`), as ONE call: every region's scan is queued on the context's stream with device-side outputs at the region's
// offset of the batch's buffers, nothing is waited for in between, and the results of all regions come home with one copy
// and one synchronisation.  Rows of region i land at index sum(n_rows[0..i)) of the outputs.
extern "C" int64_t b200_scan_regions(b200_ctx_t *c, const b200_pbf_t *pb, const b200_query_t *q, int n_regions, const int64_t *row_beg, const int64_t *n_rows,
                                     unsigned flags, b200_scan_out_t *out)
{
	if (!c || !pb || !q || !out || n_regions < 0 || (n_regions && (!row_beg || !n_rows))) { set_err("b200_scan_regions: bad argument"); return -1; }
	if (flags & B200_SCAN_DEVICE_OUT) { set_err("b200_scan_regions: outputs are host buffers"); return -1; }
	cudaSetDevice(c->dev);
	out->totals[0] = out->totals[1] = out->totals[2] = out->totals[3] = 0;
	int64_t total = 0, widest = 0;
	for (int i = 0; i < n_regions; ++i) {
		if (row_beg[i] < b200_pbf_row_beg(pb) || n_rows[i] < 0 || row_beg[i] + n_rows[i] > b200_pbf_row_end(pb)) { set_err("b200_scan_regions: region %d is not resident", i); return -1; }
		total += n_rows[i];
		if (n_rows[i] > widest) widest = n_rows[i];
	}
	if (total == 0) return 0;
	const bool host_flt = q->has_flt && q->prog.needs_host;
	const int stride = 3 + 3 * q->G, words = q->words, n_track = q->n_track;
	const bool want_counts = (flags & B200_SCAN_COUNTS) && out->counts, want_bits = (flags & B200_SCAN_HAP_BITS) && out->hap_bits[0] && out->hap_bits[1];
	const bool want_bytes = (flags & B200_SCAN_HAP_BYTES) && out->hap_bytes[0] && out->hap_bytes[1];
	if (host_flt || n_regions == 1) { // `**` filters are evaluated on the host from the counts: region by region through the ordinary scan
		int64_t done = 0;
		for (int i = 0; i < n_regions; ++i) {
			b200_scan_out_t so = *out;
			if (so.counts) so.counts += (size_t)done * stride;
			if (so.pass) so.pass += done;
			for (int p = 0; p < 2; ++p) { if (so.hap_bits[p]) so.hap_bits[p] += (size_t)done * words; if (so.hap_bytes[p]) so.hap_bytes[p] += (size_t)done * n_track; }
			if (n_rows[i] == 0) continue;
			if (b200_scan(c, pb, q, row_beg[i], n_rows[i], flags, &so) != n_rows[i]) return -1;
			for (int k = 0; k < 4; ++k) out->totals[k] += so.totals[k];
			done += n_rows[i];
		}
		return done;
	}
	const size_t nt = (size_t)total;
	bool ok = c->rg_counts.reserve(nt * stride * sizeof(int32_t) + 16) && c->rg_pass.reserve(nt + 16) &&
	          c->cnt_raw.reserve((size_t)widest * q->G * 3 * sizeof(int32_t));      // (grown before the first region is queued, not between two)
	if ((want_bits || want_bytes) && ok) for (int p = 0; p < 2 && ok; ++p) ok = c->rg_bits[p].reserve(nt * words * sizeof(uint32_t) + 16);
	if (want_bytes && ok) for (int p = 0; p < 2 && ok; ++p) ok = c->rg_bytes[p].reserve(nt * (size_t)n_track + 16);
	if (!ok) return -1;
	int64_t done = 0;
	for (int i = 0; ok && i < n_regions; ++i) {
		if (n_rows[i] == 0) continue;
		b200_scan_out_t so;
		memset(&so, 0, sizeof(so));
		so.counts = (int32_t*)c->rg_counts.p + (size_t)done * stride;
		so.pass = (uint8_t*)c->rg_pass.p + done;
		unsigned f = (flags & ~(unsigned)(B200_SCAN_HAP_BITS | B200_SCAN_HAP_BYTES)) | B200_SCAN_COUNTS | B200_SCAN_DEVICE_OUT;
		if (want_bits || want_bytes) { f |= B200_SCAN_HAP_BITS; for (int p = 0; p < 2; ++p) so.hap_bits[p] = (uint32_t*)c->rg_bits[p].p + (size_t)done * words; }
		if (want_bytes) { f |= B200_SCAN_HAP_BYTES; for (int p = 0; p < 2; ++p) so.hap_bytes[p] = (uint8_t*)c->rg_bytes[p].p + (size_t)done * n_track; }
		c->keep_totals = done > 0;
		ok = b200_scan(c, pb, q, row_beg[i], n_rows[i], f, &so) == n_rows[i];
		c->keep_totals = false;
		done += n_rows[i];
	}
	if (!ok) { cudaStreamSynchronize(c->st); return -1; }
	if (want_counts) ok = CU_OK(cudaMemcpyAsync(out->counts, c->rg_counts.p, nt * stride * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
	if (ok && out->pass) ok = CU_OK(cudaMemcpyAsync(out->pass, c->rg_pass.p, nt, cudaMemcpyDeviceToHost, c->st));
	for (int p = 0; p < 2 && ok; ++p) {
		if (want_bits) ok = CU_OK(cudaMemcpyAsync(out->hap_bits[p], c->rg_bits[p].p, nt * words * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st));
		if (ok && want_bytes) ok = CU_OK(cudaMemcpyAsync(out->hap_bytes[p], c->rg_bytes[p].p, nt * (size_t)n_track, cudaMemcpyDeviceToHost, c->st));
	}
	if (!ok) { cudaStreamSynchronize(c->st); return -1; }
	if (b200_scan_collect(c, out->totals) != 0) return -1;
	return total;
}

// after a DEVICE_OUT scan + b200_ctx_sync: fetch kernel timings and totals
extern "C" int b200_scan_collect(b200_ctx_t *c, int64_t totals[4])
{
	if (!c) return -1;
	cudaSetDevice(c->dev);
	unsigned long long dtot[4];
	int err = 0;
	bool ok = CU_OK(cudaMemcpyAsync(dtot, c->d_acc, sizeof(dtot), cudaMemcpyDeviceToHost, c->st)) &&
	          CU_OK(cudaMemcpyAsync(&err, c->d_err_scan, sizeof(err), cudaMemcpyDeviceToHost, c->st)) && CU_OK(cudaStreamSynchronize(c->st));
	if (!ok) return -1;
	if (err) { set_err("device error flags 0x%x during scan", err); return -1; }
	if (totals) for (int i = 0; i < 4; ++i) totals[i] = (int64_t)dtot[i];
	for (int i = 0; i < 4; ++i) c->last_totals[i] = (int64_t)dtot[i];
	read_scan_timings(c);
	return 0;
}

// header fields of a .pbf image without loading it (pbwt.c:231-258)
extern "C" int b200_pbf_peek(const uint8_t *f, size_t flen, int32_t *m, int32_t *shift, int64_t *n_rows)
{
	b200_pbf_t *pb = pbf_index_prepare(f, flen, 0, 0);
	if (!pb) return -1;
	if (m) *m = pb->m;
	if (shift) *shift = pb->shift;
	if (n_rows) *n_rows = pb->n;
	delete pb;
	return 0;
}

// Load + count scan in one call: the rows' pair walk, finalize and the copy of their results to the host are queued behind the
// composite maps of every chunk of the image while later chunks are still on their way, instead of after the whole load.
extern "C" b200_pbf_t *b200_pbf_load_scan(b200_ctx_t *c, const uint8_t *f, size_t flen, int64_t row_beg, int64_t row_end, const b200_query_t *q,
                                          b200_scan_out_t *out, int64_t *n_scanned)
{
	if (!c || !f || !q || !out) { set_err("b200_pbf_load_scan: null argument"); return nullptr; }
	if (q->ctx != c) { set_err("b200_pbf_load_scan: the query belongs to another context"); return nullptr; }
	cudaSetDevice(c->dev);
	int32_t m = 0; int64_t n = 0;
	if (b200_pbf_peek(f, flen, &m, nullptr, &n) != 0) return nullptr;
	if (m != q->m) { set_err("b200_pbf_load_scan: query was built for m=%d, PBF has m=%d", q->m, m); return nullptr; }
	if (row_end < 0 || row_end > n) row_end = n;
	if (row_beg < 0) row_beg = 0;
	if (row_beg > row_end) row_beg = row_end;
	out->totals[0] = out->totals[1] = out->totals[2] = out->totals[3] = 0;
	if (n_scanned) *n_scanned = row_end - row_beg;
	// what the pipeline fuses: the count-only full-cohort scan with one group and a device-side filter (BASELINE configs 2, 5);
	// any other query is the load followed by b200_scan
	const bool fuse = q->full && q->G == 1 && !(q->has_flt && q->prog.needs_host) && row_end > row_beg && getenv("BGT_B200_NO_FUSE") == nullptr;
	FusedScan fs;
	fs.q = q; fs.out = out; fs.row_lo = row_beg; fs.row_hi = row_end; fs.d_counts = nullptr; fs.d_pass = nullptr;
	b200_pbf_t *pb = pbf_load_impl(c, f, flen, row_beg, row_end, B200_LOAD_PREPARE_COUNT_SCAN, fuse ? &fs : nullptr);
	if (!pb) return nullptr;
	if (row_end == row_beg) return pb;
	bool all_split = fuse && fs.queued, none_ext = true;
	for (int b = 0; b < pb->n_blk; ++b) { all_split = all_split && pb->blk_sparse[b] == 1; none_ext = none_ext && pb->blk_sparse[b] != 2; }
	if (all_split) { // everything was queued while loading: wait for the results, fetch totals and error flags
		unsigned long long dtot[4];
		int err = 0;
		bool ok = CU_OK(cudaMemcpyAsync(dtot, c->d_acc, sizeof(dtot), cudaMemcpyDeviceToHost, c->st)) &&
		          CU_OK(cudaMemcpyAsync(&err, c->d_err_scan, sizeof(err), cudaMemcpyDeviceToHost, c->st)) &&
		          CU_OK(cudaStreamSynchronize(c->st)) && CU_OK(cudaStreamSynchronize(c->st_d2h));
		if (ok && err) { set_err("device error flags 0x%x during scan", err); ok = false; }
		if (!ok) { b200_pbf_close(pb); return nullptr; }
		for (int i = 0; i < 4; ++i) out->totals[i] = c->last_totals[i] = (int64_t)dtot[i];
		return pb;
	}
	// some block is off the split path (dense plane 1), or the query is not of the fused kind: the ordinary scan -- it keeps the
	// pair-walk counts the pipeline has already produced
	cudaStreamSynchronize(c->st_d2h);
	c->fused_pairs_done = fuse && fs.queued && none_ext;      // (a block flagged 2 was only walked as far as the pipeline's launches reach: all over again)
	const int64_t done = b200_scan(c, pb, q, row_beg, row_end - row_beg, B200_SCAN_COUNTS, out);
	c->fused_pairs_done = false;
	if (done != row_end - row_beg) { b200_pbf_close(pb); return nullptr; }
	return pb;
}

extern "C" int b200_last_totals(b200_ctx_t *c, int64_t totals[4])
{
	if (!c || !totals) return -1;
	for (int i = 0; i < 4; ++i) totals[i] = c->last_totals[i];
	return 0;
}

// ------------------------------------------------------------------------------------------------ multi-GPU totals

// The path's one collective: region shards are independent (pbwt.c:292-301), only the whole-cohort totals are summed.
// NCCL is loaded lazily so that single-GPU users (and the CLI's start-up) do not pay for it.
namespace {
struct NcclApi {
	void *h = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	bool load() {
		if (h) return true;
		const char *names[] = {"libnccl.so.2", "libnccl.so"};
		for (const char *nm : names) if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
		if (!h) { set_err("cannot load NCCL (libnccl.so.2): %s", dlerror()); return false; }
		CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
		CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
		AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
		GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
		GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
		GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
		if (!CommInitAll || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd || !GetErrorString) { set_err("libnccl.so.2 lacks an entry point"); dlclose(h); h = nullptr; return false; }
		return true;
	}
};
NcclApi g_nccl;
}

extern "C" int b200_allreduce_i64(b200_ctx_t *const *ctxs, int n, int64_t *vals, int count)
{
	if (!ctxs || !vals || n < 1 || count < 1 || n > 64) { set_err("b200_allreduce_i64: bad argument"); return -1; }
	for (int i = 0; i < n; ++i) if (!ctxs[i]) { set_err("b200_allreduce_i64: null context"); return -1; }
	if (n == 1) return 0;
	static std::mutex mu;
	std::lock_guard<std::mutex> lk(mu);
	if (!g_nccl.load()) return -1;
	std::vector<int> devs(n);
	std::vector<ncclComm_t> comms(n, nullptr);
	std::vector<int64_t*> d(n, nullptr);
	for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->dev;
	ncclResult_t r = g_nccl.CommInitAll(comms.data(), n, devs.data());
	if (r != ncclSuccess) { set_err("ncclCommInitAll over %d GPUs: %s", n, g_nccl.GetErrorString(r)); return -1; }
	bool ok = true;
	for (int i = 0; ok && i < n; ++i) {
		ok = CU_OK(cudaSetDevice(devs[i])) && CU_OK(cudaMalloc(&d[i], sizeof(int64_t) * (size_t)count)) &&
		     CU_OK(cudaMemcpyAsync(d[i], vals + (size_t)i * count, sizeof(int64_t) * (size_t)count, cudaMemcpyHostToDevice, ctxs[i]->st));
	}
	if (ok) {
		g_nccl.GroupStart();
		for (int i = 0; i < n; ++i) {
			r = g_nccl.AllReduce(d[i], d[i], (size_t)count, ncclInt64, ncclSum, comms[i], ctxs[i]->st);
			if (r != ncclSuccess) ok = false;
		}
		const ncclResult_t r2 = g_nccl.GroupEnd();
		if (r2 != ncclSuccess) { r = r2; ok = false; }
		if (!ok) set_err("ncclAllReduce: %s", g_nccl.GetErrorString(r));
	}
	for (int i = 0; ok && i < n; ++i)
		ok = CU_OK(cudaSetDevice(devs[i])) && CU_OK(cudaMemcpyAsync(vals + (size_t)i * count, d[i], sizeof(int64_t) * (size_t)count, cudaMemcpyDeviceToHost, ctxs[i]->st)) &&
		     CU_OK(cudaStreamSynchronize(ctxs[i]->st));
	for (int i = 0; i < n; ++i) { cudaSetDevice(devs[i]); if (d[i]) cudaFree(d[i]); if (comms[i]) g_nccl.CommDestroy(comms[i]); ++ctxs[i]->launches; }
	return ok ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------ synthetic cohort

extern "C" b200_pbf_t *b200_synth_generate(b200_ctx_t *c, const b200_synth_t *cfg)
{
	if (!c || !cfg) { set_err("b200_synth_generate: null argument"); return nullptr; }
	cudaSetDevice(c->dev);
	if (cfg->n_samples <= 0 || cfg->n_samples >= (1 << 29) || cfg->n_rows <= 0 || cfg->shift < 0 || cfg->shift > 24) { set_err("b200_synth_generate: bad shape"); return nullptr; }
	SynthCfg sc;
	sc.m = 2u * (uint32_t)cfg->n_samples; sc.n_rows = cfg->n_rows; sc.shift = cfg->shift; sc.seed = cfg->seed;
	sc.r_max = cfg->r_max > 0 ? cfg->r_max : 64; sc.p1_one_in = cfg->p1_one_in > 0 ? cfg->p1_one_in : 16;
	sc.p1_max_iv = cfg->p1_max_iv > 0 ? cfg->p1_max_iv : 3; sc.p1_max_len = cfg->p1_max_len > 0 ? cfg->p1_max_len : 64;
	const int BS = 1 << sc.shift;
	const int64_t n = sc.n_rows;
	const int nb = (int)((n + BS - 1) / BS);
	const size_t snap = 1 + 8 * (size_t)sc.m;

	b200_pbf_t *pb = new b200_pbf_t();
	pb->ctx = c; pb->m = (int)sc.m; pb->g = 2; pb->shift = sc.shift; pb->BS = BS; pb->n = n; pb->n_blk_file = nb; pb->blk0 = 0; pb->n_blk = nb;
	uint32_t *d_len2 = nullptr; uint64_t *d_flat = nullptr;
	std::vector<uint32_t> len2((size_t)n * 2);
	std::vector<uint64_t> flat((size_t)n);
	std::vector<uint8_t> tail;
	bool ok = CU_OK(cudaMalloc(&d_len2, sizeof(uint32_t) * 2 * (size_t)n));
	ok = ok && CU_OK(launch_synth_lengths(sc, d_len2, c->st));
	++c->launches;
	ok = ok && CU_OK(cudaMemcpyAsync(len2.data(), d_len2, sizeof(uint32_t) * 2 * (size_t)n, cudaMemcpyDeviceToHost, c->st)) && CU_OK(cudaStreamSynchronize(c->st));
	if (d_len2) cudaFree(d_len2);
	if (!ok) { delete pb; return nullptr; }
	// file layout: header, per block 'S' + 2 planes of int32[m], rows, 'I' record, trailing offset
	pb->rows_in_blk.resize(nb);
	pb->h_rowoff.assign((size_t)nb * (BS + 1), 0);
	pb->h_blkoff.resize(nb);
	pb->h_blkend.resize(nb);
	uint64_t pos = 16;
	for (int b = 0; b < nb; ++b) {
		const int64_t r0 = (int64_t)b << sc.shift;
		const int rows = (int)(n - r0 < BS ? n - r0 : BS);
		pb->rows_in_blk[b] = rows;
		pb->h_blkoff[b] = pos;
		pos += snap;
		uint64_t *ro = pb->h_rowoff.data() + (size_t)b * (BS + 1);
		for (int r = 0; r < rows; ++r) {
			ro[r] = flat[r0 + r] = pos;
			pos += 9 + (uint64_t)len2[(r0 + r) * 2] + len2[(r0 + r) * 2 + 1];
		}
		ro[rows] = pos;
		pb->h_blkend[b] = pos;
	}
	const uint64_t ioff = pos;
	tail.resize(1 + 8 + 4 + 8 * (size_t)nb + 8);
	{ // pbwt.c:268-276
		uint8_t *t = tail.data();
		const int32_t n_idx = nb;
		*t++ = 'I'; memcpy(t, &n, 8); t += 8; memcpy(t, &n_idx, 4); t += 4;
		memcpy(t, pb->h_blkoff.data(), 8 * (size_t)nb); t += 8 * (size_t)nb;
		memcpy(t, &ioff, 8);
	}
	pb->file_size = pb->img_bytes = (size_t)(ioff + tail.size());
	pb->file_off0 = 0;
	pb->h_idx = pb->h_blkoff; pb->ioff = ioff;
	uint8_t hdr[16];
	{ const int32_t v[3] = {(int32_t)sc.m, 2, sc.shift}; memcpy(hdr, "PBF\1", 4); memcpy(hdr + 4, v, 12); } // pbwt.c:214-216
	ok = pool_malloc(c, (void**)&pb->d_img, pb->img_bytes + 64) && CU_OK(cudaMalloc(&d_flat, sizeof(uint64_t) * (size_t)n));
	ok = ok && CU_OK(cudaMemsetAsync(pb->d_img, 0, pb->img_bytes + 64, c->st));
	ok = ok && CU_OK(cudaMemcpyAsync(pb->d_img, hdr, 16, cudaMemcpyHostToDevice, c->st));
	ok = ok && CU_OK(cudaMemcpyAsync(pb->d_img + ioff, tail.data(), tail.size(), cudaMemcpyHostToDevice, c->st));
	for (int b = 0; ok && b < nb; ++b) ok = CU_OK(cudaMemsetAsync(pb->d_img + pb->h_blkoff[b], 'S', 1, c->st));
	ok = ok && CU_OK(cudaMemcpyAsync(d_flat, flat.data(), sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, c->st));
	ok = ok && CU_OK(launch_synth_write(sc, d_flat, pb->d_img, c->st));
	++c->launches;
	ok = ok && CU_OK(cudaStreamSynchronize(c->st));
	if (d_flat) cudaFree(d_flat);
	// row index, tiles and n1 on the device (index.cu), then the running permutation: the walk kernel in CHAIN mode over
	// all rows (identity start, pbwt.c:103) dumps S = rank^-1 in front of every block; then start ranks + plane-1 view
	ok = ok && pbf_alloc_index(pb, c->st);
	ok = ok && CU_OK(launch_index(index_params(pb, 0), nb, c->st));
	++c->launches;
	ok = ok && queue_tiles_rowmeta(pb, 0, nb, c->st);
	if (ok) {
		WalkParams P;
		memset(&P, 0, sizeof(P));
		P.img = pb->d_img; P.rowoff = pb->d_rowoff; P.n1 = pb->d_n1; P.tiles = pb->d_tiles; P.blk_tile_beg = pb->d_blk_tile_beg; P.blk_tile_end = pb->d_blk_tile_end;
		P.snap_img = pb->d_img; P.blkoff = pb->d_blkoff;
		P.m = pb->m; P.n_track = pb->m; P.G = 1; P.words = (pb->m + 31) / 32; P.shift = pb->shift;
		P.n_blk_chain = nb; P.blk_row0 = 0; P.row_lo = 0; P.row_hi = pb->n; P.err = c->d_err;
		uint8_t *d_zero = nullptr;   // group map of the (unused) counters
		ok = CU_OK(cudaMalloc(&d_zero, (size_t)pb->m + 16)) && CU_OK(cudaMemsetAsync(d_zero, 0, (size_t)pb->m + 16, c->st));
		P.tgrp = d_zero;
		const int C = pb->m > 148 * 2 * WALK_NT * 4 ? 4 : pb->m > 148 * 2 * WALK_NT * 2 ? 2 : 1;
		const int slices = (pb->m + WALK_NT * C - 1) / (WALK_NT * C);
		ok = ok && CU_OK(launch_walk(P, C, WALK_MODE_CHAIN, slices, nb, c->st));
		++c->launches;
		ok = ok && CU_OK(cudaStreamSynchronize(c->st));
		if (d_zero) cudaFree(d_zero);
	}
	ok = ok && queue_ranks_view(pb, 0, nb, c->st) && pbf_finish_load(pb);
	if (!ok) { cudaStreamSynchronize(c->st); pbf_free_device(pb); delete pb; return nullptr; }
	return pb;
}

// ------------------------------------------------------------------------------------------------ encoder (pbf_open_w / pbf_write / pbf_close)

struct b200_enc_s {
	b200_ctx_t *ctx = nullptr;
	int m = 0, shift = 0, words = 0;
	int64_t n = 0;                       // rows written so far (pbf_t.n, pbwt.c:309)
	int32_t *d_rank = nullptr;
	uint32_t *d_bitvec = nullptr;
	unsigned long long *d_pos = nullptr; // [4]: stream offsets in / out
	DevBuf in_bits, bytes[2], snap, out[2], row_len;
	std::vector<uint8_t> image;          // the file so far, minus what b200_enc_drain has handed out
	std::vector<uint8_t> drained;        // the bytes of the last b200_enc_drain (valid until the next call on this encoder)
	uint64_t file_off = 0;               // file offset of image[0] = bytes drained so far
	std::vector<uint64_t> idx;           // offsets of the 'S' records (pbwt.c:297)
	std::vector<uint8_t> h_out[2];
	std::vector<uint32_t> h_len;
	std::vector<int32_t> h_snap;
	bool finished = false;
	int batch_rows = 0;
};

extern "C" void b200_enc_destroy(b200_enc_t *e)
{
	if (!e) return;
	cudaSetDevice(e->ctx->dev);
	cudaStreamSynchronize(e->ctx->st);
	if (e->d_rank) cudaFree(e->d_rank);
	if (e->d_bitvec) cudaFree(e->d_bitvec);
	if (e->d_pos) cudaFree(e->d_pos);
	e->in_bits.release(); e->bytes[0].release(); e->bytes[1].release(); e->snap.release(); e->out[0].release(); e->out[1].release(); e->row_len.release();
	delete e;
}

extern "C" b200_enc_t *b200_enc_create(b200_ctx_t *c, int m, int shift)
{
	if (!c) { set_err("b200_enc_create: null context"); return nullptr; }
	if (m <= 0 || m >= (1 << 30) || shift < 0 || shift > 24) { set_err("b200_enc_create: m=%d shift=%d out of range", m, shift); return nullptr; }
	if (encode_smem_bytes(m) > 220 * 1024) { set_err("b200_enc_create: m=%d needs more shared memory than one SM has (limit about 1.2 M columns)", m); return nullptr; }
	cudaSetDevice(c->dev);
	b200_enc_t *e = new b200_enc_t();
	e->ctx = c; e->m = m; e->shift = shift; e->words = (m + 31) / 32;
	long long br = (256LL << 20) / m;
	e->batch_rows = (int)(br < 16 ? 16 : (br > 4096 ? 4096 : br));
	std::vector<int32_t> ident(2 * (size_t)m);
	for (int p = 0; p < 2; ++p) for (int i = 0; i < m; ++i) ident[(size_t)p * m + i] = i;   // pbwt.c:103: S starts as the identity
	bool ok = CU_OK(cudaMalloc(&e->d_rank, sizeof(int32_t) * 2 * (size_t)m)) && CU_OK(cudaMalloc(&e->d_bitvec, sizeof(uint32_t) * 6 * (size_t)e->words + 16)) &&
	          CU_OK(cudaMalloc(&e->d_pos, 4 * sizeof(unsigned long long)));
	ok = ok && CU_OK(cudaMemcpyAsync(e->d_rank, ident.data(), sizeof(int32_t) * 2 * (size_t)m, cudaMemcpyHostToDevice, c->st)) &&
	     CU_OK(cudaMemsetAsync(e->d_bitvec, 0, sizeof(uint32_t) * 6 * (size_t)e->words, c->st)) && CU_OK(cudaStreamSynchronize(c->st));
	if (!ok) { b200_enc_destroy(e); return nullptr; }
	// file header (pbwt.c:214-216)
	e->image.resize(16);
	const int32_t v[3] = {m, 2, shift};
	memcpy(e->image.data(), "PBF\1", 4);
	memcpy(e->image.data() + 4, v, 12);
	return e;
}

// rows [e->n, e->n + R) are in e->in_bits (device): encode them and append their records to the image
static bool enc_batch(b200_enc_t *e, int R)
{
	b200_ctx_t *c = e->ctx;
	const int m = e->m, words = e->words;
	const int64_t BS = 1LL << e->shift, row0 = e->n;
	const int64_t k0 = (row0 + BS - 1) >> e->shift, k1 = (row0 + R - 1) >> e->shift;   // checkpoints inside the batch: k0..k1
	const int n_snap = (int)(k1 >= k0 ? k1 - k0 + 1 : 0);
	if (!e->snap.reserve(sizeof(int32_t) * 2 * (size_t)m * (size_t)(n_snap ? n_snap : 1)) || !e->row_len.reserve(sizeof(uint32_t) * 2 * (size_t)R)) return false;
	for (int p = 0; p < 2; ++p) if (!e->out[p].reserve((size_t)R * (size_t)m + 64)) return false;   // a row's code never exceeds m bytes per plane
	EncodeParams P;
	memset(&P, 0, sizeof(P));
	P.in_bits = (const uint32_t*)e->in_bits.p; P.m = m; P.words = words; P.shift = e->shift; P.n_rows = R; P.row0 = row0;
	P.rank = e->d_rank; P.bitvec = e->d_bitvec; P.snap = (int32_t*)e->snap.p; P.out[0] = (uint8_t*)e->out[0].p; P.out[1] = (uint8_t*)e->out[1].p;
	P.row_len = (uint32_t*)e->row_len.p; P.out_pos0 = e->d_pos; P.out_pos1 = e->d_pos + 2;
	unsigned long long pos[2] = {0, 0};
	bool ok = CU_OK(cudaMemsetAsync(e->d_pos, 0, 4 * sizeof(unsigned long long), c->st)) && CU_OK(launch_encode(P, c->sm_count, c->st));
	++c->launches;
	e->h_len.resize(2 * (size_t)R);
	ok = ok && CU_OK(cudaMemcpyAsync(pos, e->d_pos + 2, sizeof(pos), cudaMemcpyDeviceToHost, c->st)) &&
	     CU_OK(cudaMemcpyAsync(e->h_len.data(), e->row_len.p, sizeof(uint32_t) * 2 * (size_t)R, cudaMemcpyDeviceToHost, c->st)) &&
	     CU_OK(cudaStreamSynchronize(c->st));
	if (!ok) return false;
	for (int p = 0; p < 2; ++p) {
		e->h_out[p].resize((size_t)pos[p] + 1);
		if (pos[p] && !CU_OK(cudaMemcpyAsync(e->h_out[p].data(), e->out[p].p, (size_t)pos[p], cudaMemcpyDeviceToHost, c->st))) return false;
	}
	e->h_snap.resize(2 * (size_t)m * (size_t)(n_snap ? n_snap : 1));
	if (n_snap && !CU_OK(cudaMemcpyAsync(e->h_snap.data(), e->snap.p, sizeof(int32_t) * 2 * (size_t)m * n_snap, cudaMemcpyDeviceToHost, c->st))) return false;
	if (!CU_OK(cudaStreamSynchronize(c->st))) return false;
	// file assembly, as pbf_write does it row by row (pbwt.c:288-311)
	size_t o[2] = {0, 0};
	for (int r = 0; r < R; ++r) {
		const int64_t arow = row0 + r;
		if ((arow & (BS - 1)) == 0) {
			e->idx.push_back(e->file_off + (uint64_t)e->image.size());
			e->image.push_back('S');
			const uint8_t *s = (const uint8_t*)(e->h_snap.data() + 2 * (size_t)m * (size_t)((arow >> e->shift) - k0));
			e->image.insert(e->image.end(), s, s + 8 * (size_t)m);
		}
		e->image.push_back('B');
		for (int p = 0; p < 2; ++p) {
			const uint32_t l = e->h_len[2 * (size_t)r + p];
			const uint8_t *lb = (const uint8_t*)&l;
			e->image.insert(e->image.end(), lb, lb + 4);
			e->image.insert(e->image.end(), e->h_out[p].data() + o[p], e->h_out[p].data() + o[p] + l);
			o[p] += l;
		}
	}
	e->n += R;
	return true;
}

extern "C" int b200_enc_write_bits(b200_enc_t *e, const uint32_t *bits, int64_t n_rows)
{
	if (!e || (!bits && n_rows > 0) || n_rows < 0) { set_err("b200_enc_write_bits: bad argument"); return -1; }
	if (e->finished) { set_err("b200_enc_write_bits: the image was already finished"); return -1; }
	cudaSetDevice(e->ctx->dev);
	const size_t row_words = 2 * (size_t)e->words;
	for (int64_t done = 0; done < n_rows;) {
		const int R = (int)(n_rows - done < e->batch_rows ? n_rows - done : e->batch_rows);
		if (!e->in_bits.reserve(sizeof(uint32_t) * row_words * (size_t)R)) return -1;
		if (!CU_OK(cudaMemcpyAsync(e->in_bits.p, bits + (size_t)done * row_words, sizeof(uint32_t) * row_words * (size_t)R, cudaMemcpyHostToDevice, e->ctx->st))) return -1;
		if (!enc_batch(e, R)) return -1;
		done += R;
	}
	return 0;
}

extern "C" int b200_enc_write_bytes(b200_enc_t *e, const uint8_t *a0, const uint8_t *a1, int64_t n_rows)
{
	if (!e || ((!a0 || !a1) && n_rows > 0) || n_rows < 0) { set_err("b200_enc_write_bytes: bad argument"); return -1; }
	if (e->finished) { set_err("b200_enc_write_bytes: the image was already finished"); return -1; }
	cudaSetDevice(e->ctx->dev);
	const size_t m = (size_t)e->m, row_words = 2 * (size_t)e->words;
	for (int64_t done = 0; done < n_rows;) {
		const int R = (int)(n_rows - done < e->batch_rows ? n_rows - done : e->batch_rows);
		if (!e->in_bits.reserve(sizeof(uint32_t) * row_words * (size_t)R) || !e->bytes[0].reserve(m * R) || !e->bytes[1].reserve(m * R)) return -1;
		bool ok = CU_OK(cudaMemcpyAsync(e->bytes[0].p, a0 + (size_t)done * m, m * R, cudaMemcpyHostToDevice, e->ctx->st)) &&
		          CU_OK(cudaMemcpyAsync(e->bytes[1].p, a1 + (size_t)done * m, m * R, cudaMemcpyHostToDevice, e->ctx->st)) &&
		          CU_OK(launch_pack_rows((const uint8_t*)e->bytes[0].p, (const uint8_t*)e->bytes[1].p, R, e->m, (uint32_t*)e->in_bits.p, e->ctx->st));
		++e->ctx->launches;
		if (!ok || !enc_batch(e, R)) return -1;
		done += R;
	}
	return 0;
}

extern "C" int64_t b200_enc_rows(const b200_enc_t *e) { return e ? e->n : -1; }

extern "C" int64_t b200_enc_finish(b200_enc_t *e, const uint8_t **image)
{
	if (!e) { set_err("b200_enc_finish: null encoder"); return -1; }
	if (!e->finished) { // the index record (pbwt.c:268-276)
		const uint64_t off = e->file_off + (uint64_t)e->image.size();
		const int64_t n = e->n;
		const int32_t n_idx = (int32_t)e->idx.size();
		e->image.push_back('I');
		e->image.insert(e->image.end(), (const uint8_t*)&n, (const uint8_t*)&n + 8);
		e->image.insert(e->image.end(), (const uint8_t*)&n_idx, (const uint8_t*)&n_idx + 4);
		e->image.insert(e->image.end(), (const uint8_t*)e->idx.data(), (const uint8_t*)e->idx.data() + 8 * (size_t)n_idx);
		e->image.insert(e->image.end(), (const uint8_t*)&off, (const uint8_t*)&off + 8);
		e->finished = true;
	}
	if (image) *image = e->image.data();
	return (int64_t)e->image.size();
}

// Streaming writers (pbf_write writes every row as it goes, pbwt.c:288-311): hand out the bytes assembled since the last
// drain and forget them; only the block index and the running file offset are kept for the 'I' record.
extern "C" int64_t b200_enc_drain(b200_enc_t *e, const uint8_t **bytes)
{
	if (!e || !bytes) { set_err("b200_enc_drain: null argument"); return -1; }
	e->drained.clear();
	e->drained.swap(e->image);
	e->file_off += (uint64_t)e->drained.size();
	*bytes = e->drained.data();
	return (int64_t)e->drained.size();
}

// ------------------------------------------------------------------------------------------------ BGZF (bgzf.c)

struct BgzfIndex { std::vector<uint64_t> coff, uoff; std::vector<uint32_t> csize, usize; uint64_t total = 0; uint32_t max_csize = 0; };

// walk the block headers of a BGZF image (bgzf.c:259-281 check_header, :318-351 bgzf_read_block); ISIZE gives the output layout
static bool bgzf_index(const uint8_t *f, size_t n, BgzfIndex &ix)
{
	size_t pos = 0;
	while (pos < n) {
		if (pos + 18 > n || f[pos] != 31 || f[pos + 1] != 139 || f[pos + 2] != 8 || !(f[pos + 3] & 4)) { set_err("not a BGZF file (block header at offset %zu)", pos); return false; }
		const uint32_t xlen = f[pos + 10] | (uint32_t)f[pos + 11] << 8;
		if (pos + 12 + xlen > n) { set_err("truncated BGZF block header"); return false; }
		uint32_t bsize = 0; bool found = false;
		for (uint32_t x = 0; x + 4 <= xlen;) { // extra subfields: SI1 SI2 SLEN data
			const uint8_t *e = f + pos + 12 + x;
			const uint32_t slen = e[2] | (uint32_t)e[3] << 8;
			if (e[0] == 'B' && e[1] == 'C' && slen == 2 && x + 6 <= xlen) { bsize = (e[4] | (uint32_t)e[5] << 8) + 1; found = true; }
			x += 4 + slen;
		}
		if (!found || bsize < 12 + xlen + 8 || pos + bsize > n) { set_err("corrupt BGZF block at offset %zu", pos); return false; }
		const uint32_t csz = bsize - 12 - xlen - 8;
		uint32_t isize;
		memcpy(&isize, f + pos + bsize - 4, 4);
		if (isize > 65536) { set_err("BGZF block at offset %zu claims %u uncompressed bytes", pos, isize); return false; }
		if (isize) { // the empty EOF block (bgzf.c:51-57) carries nothing
			ix.coff.push_back((uint64_t)pos + 12 + xlen); ix.csize.push_back(csz); ix.usize.push_back(isize); ix.uoff.push_back(ix.total);
			ix.total += isize;
			if (csz > ix.max_csize) ix.max_csize = csz;
		}
		pos += bsize;
	}
	return true;
}

// inflate a BGZF image into a device buffer (from the pool; caller frees with pool_free).  No sync.
static bool bgzf_inflate_device(b200_ctx_t *c, const uint8_t *f, size_t n, uint8_t **d_out, uint64_t *out_len)
{
	BgzfIndex ix;
	if (!bgzf_index(f, n, ix)) return false;
	const size_t nb = ix.coff.size();
	uint8_t *d_in = nullptr, *d_o = nullptr, *d_tab = nullptr;
	const size_t tab_bytes = nb * (8 + 8 + 4 + 4) + 64;
	bool ok = pool_malloc(c, (void**)&d_in, n + 64) && pool_malloc(c, (void**)&d_o, (size_t)ix.total + 64) && pool_malloc(c, (void**)&d_tab, tab_bytes);
	if (!ok) return false;
	uint64_t *d_coff = (uint64_t*)d_tab, *d_uoff = d_coff + nb;
	uint32_t *d_csize = (uint32_t*)(d_uoff + nb), *d_usize = d_csize + nb;
	ok = CU_OK(cudaMemcpyAsync(d_in, f, n, cudaMemcpyHostToDevice, c->st));
	if (nb) ok = ok && CU_OK(cudaMemcpyAsync(d_coff, ix.coff.data(), nb * 8, cudaMemcpyHostToDevice, c->st)) && CU_OK(cudaMemcpyAsync(d_uoff, ix.uoff.data(), nb * 8, cudaMemcpyHostToDevice, c->st)) &&
	     CU_OK(cudaMemcpyAsync(d_csize, ix.csize.data(), nb * 4, cudaMemcpyHostToDevice, c->st)) && CU_OK(cudaMemcpyAsync(d_usize, ix.usize.data(), nb * 4, cudaMemcpyHostToDevice, c->st));
	InflateParams P;
	memset(&P, 0, sizeof(P));
	P.in = d_in; P.blk_coff = d_coff; P.blk_csize = d_csize; P.blk_usize = d_usize; P.blk_uoff = d_uoff; P.out = d_o; P.max_csize = ix.max_csize; P.err = c->d_err_sites;
	ok = ok && CU_OK(cudaMemsetAsync(c->d_err_sites, 0, sizeof(int), c->st)) && CU_OK(launch_bgzf_inflate(P, (int)nb, c->st));
	++c->launches;
	// the tables and the compressed image are read by the queued kernel: released after the stream has passed it
	ok = ok && CU_OK(cudaStreamSynchronize(c->st));
	pool_free(c, d_in); pool_free(c, d_tab);
	if (!ok) { pool_free(c, d_o); return false; }
	int err = 0;
	if (!CU_OK(cudaMemcpy(&err, c->d_err_sites, sizeof(int), cudaMemcpyDeviceToHost))) { pool_free(c, d_o); return false; }
	if (err & 256) { set_err("corrupt BGZF file: a block does not inflate to its ISIZE"); pool_free(c, d_o); return false; }
	*d_out = d_o; *out_len = ix.total;
	return true;
}

extern "C" int64_t b200_bgzf_inflate(b200_ctx_t *c, const uint8_t *bytes, size_t n_bytes, uint8_t *out, size_t out_cap)
{
	if (!bytes) { set_err("b200_bgzf_inflate: null argument"); return -1; }
	if (!out) { // size query: block headers only, host logic (no device needed)
		BgzfIndex ix;
		return bgzf_index(bytes, n_bytes, ix) ? (int64_t)ix.total : -1;
	}
	if (!c) { set_err("b200_bgzf_inflate: null context"); return -1; }
	cudaSetDevice(c->dev);
	uint8_t *d = nullptr; uint64_t len = 0;
	if (!bgzf_inflate_device(c, bytes, n_bytes, &d, &len)) return -1;
	bool ok = true;
	if (len > out_cap) { set_err("b200_bgzf_inflate: output needs %llu bytes", (unsigned long long)len); ok = false; }
	ok = ok && (len == 0 || (CU_OK(cudaMemcpyAsync(out, d, (size_t)len, cudaMemcpyDeviceToHost, c->st)) && CU_OK(cudaStreamSynchronize(c->st))));
	pool_free(c, d);
	return ok ? (int64_t)len : -1;
}

// ------------------------------------------------------------------------------------------------ sites (.bcf / .csi) and the text of `view -G`

struct b200_sites_s {
	b200_ctx_t *ctx = nullptr;
	uint8_t *d_bcf = nullptr; uint64_t bcf_len = 0;
	SiteRec *d_sites = nullptr;
	int64_t n = 0;
	std::string header;                 // BCF header text (vcf.c:263-288)
	std::vector<std::string> contigs;   // BCF_DT_CTG in id order (vcf.c:121-133)
	char *d_ctg = nullptr; int *d_ctg_off = nullptr;
	int row_key = -1;
	std::vector<int64_t> h_rows;        // INFO/_row of every record (host copy), and whether they ascend (records in row order)
	bool rows_sorted = true;
	DevBuf len, off, temp, text;
	char *h_text = nullptr; size_t h_text_cap = 0;
	unsigned long long *d_nlines = nullptr;
};

extern "C" void b200_sites_destroy(b200_sites_t *s)
{
	if (!s) return;
	cudaSetDevice(s->ctx->dev);
	cudaStreamSynchronize(s->ctx->st);
	pool_free(s->ctx, s->d_bcf);
	if (s->d_sites) cudaFree(s->d_sites);
	if (s->d_ctg) cudaFree(s->d_ctg);
	if (s->d_ctg_off) cudaFree(s->d_ctg_off);
	if (s->d_nlines) cudaFree(s->d_nlines);
	s->len.release(); s->off.release(); s->temp.release(); s->text.release();
	if (s->h_text) cudaFreeHost(s->h_text);
	delete s;
}

// ID of a "##KIND=<...ID=xxx...>" header line, or empty
static std::string hdr_line_id(const std::string &ln)
{
	size_t p = ln.find('<');
	if (p == std::string::npos) return "";
	size_t q = ln.find("ID=", p);
	while (q != std::string::npos && !(ln[q - 1] == '<' || ln[q - 1] == ',')) q = ln.find("ID=", q + 1);
	if (q == std::string::npos) return "";
	size_t e = ln.find_first_of(",>", q + 3);
	return ln.substr(q + 3, e == std::string::npos ? std::string::npos : e - q - 3);
}

// record-number index at the tail of an inflated .csi (hts.c:536-542): "RNI\1", n_rec, rec_shift, n, voff[n]
static bool csi_rni(const std::vector<uint8_t> &csi, uint64_t *n_rec, int *rec_shift, std::vector<uint64_t> &voff)
{
	if (csi.size() < 16 || memcmp(csi.data(), "CSI\1", 4) != 0) return false;
	size_t p = 4;
	int32_t v[3];
	memcpy(v, csi.data() + p, 12); p += 12;                       // min_shift, n_lvls, l_meta (hts.c:529-533)
	if (v[2] < 0 || p + (size_t)v[2] + 4 > csi.size()) return false;
	p += (size_t)v[2];
	int32_t n_ref;
	memcpy(&n_ref, csi.data() + p, 4); p += 4;
	for (int32_t r = 0; r < n_ref; ++r) {                         // hts_idx_save_core, CSI flavour (hts.c:470-510)
		int32_t n_bin;
		if (p + 4 > csi.size()) return false;
		memcpy(&n_bin, csi.data() + p, 4); p += 4;
		for (int32_t b = 0; b < n_bin; ++b) {
			int32_t n_chunk;
			if (p + 16 > csi.size()) return false;
			memcpy(&n_chunk, csi.data() + p + 12, 4);             // bin u32, loff u64, n_chunk i32
			p += 16;
			if (n_chunk < 0 || p + 16ull * (size_t)n_chunk > csi.size()) return false;
			p += 16ull * (size_t)n_chunk;
		}
	}
	p += 8;                                                       // n_no_coor
	if (p + 20 > csi.size() || memcmp(csi.data() + p, "RNI\1", 4) != 0) return false;
	int32_t n;
	memcpy(n_rec, csi.data() + p + 4, 8); memcpy(rec_shift, csi.data() + p + 12, 4); memcpy(&n, csi.data() + p + 16, 4);
	p += 20;
	if (n < 0 || *rec_shift <= 0 || *rec_shift > 30 || p + 8ull * (size_t)n > csi.size()) return false;
	voff.resize((size_t)n);
	memcpy(voff.data(), csi.data() + p, 8ull * (size_t)n);
	return true;
}

extern "C" b200_sites_t *b200_sites_load(b200_ctx_t *c, const uint8_t *bcf, size_t n_bcf, const uint8_t *csi, size_t n_csi, int row_key)
{
	if (!c || !bcf) { set_err("b200_sites_load: null argument"); return nullptr; }
	cudaSetDevice(c->dev);
	b200_sites_t *s = new b200_sites_t();
	s->ctx = c;
	BgzfIndex ix;
	if (!bgzf_index(bcf, n_bcf, ix) || !bgzf_inflate_device(c, bcf, n_bcf, &s->d_bcf, &s->bcf_len)) { b200_sites_destroy(s); return nullptr; }
	// ---- header: "BCF\2\2", l_text, text (vcf.c:263-288)
	uint8_t h9[9];
	uint32_t l_text = 0;
	bool ok = s->bcf_len >= 9 && CU_OK(cudaMemcpy(h9, s->d_bcf, 9, cudaMemcpyDeviceToHost)) && memcmp(h9, "BCF\2\2", 5) == 0;
	if (ok) { memcpy(&l_text, h9 + 5, 4); ok = 9ull + l_text <= s->bcf_len; }
	if (!ok) { set_err("not a BCF2 file"); b200_sites_destroy(s); return nullptr; }
	s->header.resize(l_text);
	if (l_text && !CU_OK(cudaMemcpy(&s->header[0], s->d_bcf + 9, l_text, cudaMemcpyDeviceToHost))) { b200_sites_destroy(s); return nullptr; }
	while (!s->header.empty() && s->header.back() == 0) s->header.pop_back();
	{ // dictionaries as bcf_hdr_parse builds them (vcf.c:193-208): PASS first, then IDs / contigs in order of first appearance
		std::vector<std::string> ids(1, "PASS");
		size_t a = 0;
		while (a < s->header.size()) {
			size_t e = s->header.find('\n', a);
			if (e == std::string::npos) e = s->header.size();
			const std::string ln = s->header.substr(a, e - a);
			a = e + 1;
			if (ln.compare(0, 2, "##") != 0) continue;
			const std::string id = hdr_line_id(ln);
			if (id.empty()) continue;
			if (ln.compare(0, 9, "##contig=") == 0) { if (std::find(s->contigs.begin(), s->contigs.end(), id) == s->contigs.end()) s->contigs.push_back(id); }
			else if (ln.compare(0, 9, "##FILTER=") == 0 || ln.compare(0, 7, "##INFO=") == 0 || ln.compare(0, 9, "##FORMAT=") == 0) {
				if (std::find(ids.begin(), ids.end(), id) == ids.end()) ids.push_back(id);
			}
		}
		if (row_key < 0) { for (size_t i = 0; i < ids.size(); ++i) if (ids[i] == "_row") row_key = (int)i; }
		s->row_key = row_key;
	}
	// ---- record offsets: stretches from the RNI of the .csi, or one stretch from the first record
	const uint64_t first = 9ull + l_text;
	std::vector<unsigned long long> seg;
	int seg_len = 0;
	int64_t n_rec = -1;
	if (csi && n_csi) {
		const int64_t clen = b200_bgzf_inflate(c, csi, n_csi, nullptr, 0);
		std::vector<uint8_t> raw(clen > 0 ? (size_t)clen : 1);
		std::vector<uint64_t> voff;
		uint64_t nr = 0; int shift = 0;
		if (clen > 0 && b200_bgzf_inflate(c, csi, n_csi, raw.data(), raw.size()) == clen && (raw.resize((size_t)clen), csi_rni(raw, &nr, &shift, voff))) {
			// virtual offset = file offset of the BGZF block << 16 | offset inside it (bgzf.h); blocks are in ix by their DEFLATE start
			std::vector<uint64_t> blk_start(ix.coff.size());
			for (size_t b = 0; b < ix.coff.size(); ++b) blk_start[b] = ix.coff[b];
			bool good = (uint64_t)voff.size() == ((nr + (1ull << shift) - 1) >> shift);
			for (size_t k = 0; good && k < voff.size(); ++k) {
				const uint64_t co = voff[k] >> 16, uo = voff[k] & 0xffff;
				// the block whose header starts at co: its DEFLATE stream starts 12 + xlen bytes later -> last block with start > co is the one behind it
				size_t b = std::upper_bound(blk_start.begin(), blk_start.end(), co) - blk_start.begin();
				if (b >= blk_start.size() || blk_start[b] - co > 12 + 65535) { // co may point at the EOF block / end of file: position = end of stream
					if (co >= n_bcf - 28 || b >= blk_start.size()) { seg.push_back(s->bcf_len); continue; }
					good = false; break;
				}
				seg.push_back(ix.uoff[b] + uo);
			}
			if (good) { n_rec = (int64_t)nr; seg_len = 1 << shift; } else seg.clear();
		}
	}
	unsigned long long *d_seg = nullptr, *d_off = nullptr, *d_cnt = nullptr;
	ok = CU_OK(cudaMemsetAsync(c->d_err_sites, 0, sizeof(int), c->st));
	if (n_rec < 0) { // no usable RNI: count, then chase, from the first record with one thread
		seg.assign(1, first);
		ok = ok && CU_OK(cudaMalloc(&d_seg, 8)) && CU_OK(cudaMalloc(&d_cnt, 8)) && CU_OK(cudaMemcpyAsync(d_seg, seg.data(), 8, cudaMemcpyHostToDevice, c->st)) &&
		     CU_OK(launch_bcf_chase(s->d_bcf, s->bcf_len, d_seg, 1, 0, 0, nullptr, d_cnt, c->d_err_sites, c->st));
		unsigned long long cnt = 0;
		ok = ok && CU_OK(cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, c->st)) && CU_OK(cudaStreamSynchronize(c->st));
		++c->launches;
		n_rec = (int64_t)cnt; seg_len = n_rec > 0 ? (int)(n_rec < (1LL << 30) ? n_rec : (1LL << 30)) : 1;
	} else {
		ok = ok && CU_OK(cudaMalloc(&d_seg, 8 * (seg.size() + 1))) && CU_OK(cudaMemcpyAsync(d_seg, seg.data(), 8 * seg.size(), cudaMemcpyHostToDevice, c->st));
	}
	s->n = n_rec;
	ok = ok && CU_OK(cudaMalloc(&d_off, 8 * (size_t)(n_rec + 1))) && CU_OK(cudaMalloc(&s->d_sites, sizeof(SiteRec) * (size_t)(n_rec + 1)));
	ok = ok && CU_OK(launch_bcf_chase(s->d_bcf, s->bcf_len, d_seg, (int)seg.size(), seg_len, n_rec, d_off, nullptr, c->d_err_sites, c->st)) &&
	     CU_OK(launch_bcf_parse(s->d_bcf, s->bcf_len, d_off, n_rec, s->row_key, s->d_sites, c->d_err_sites, c->st));
	c->launches += 2;
	// contig names for the text kernels
	std::string names; std::vector<int> coff(1, 0);
	for (const std::string &nm : s->contigs) { names += nm; coff.push_back((int)names.size()); }
	if (s->contigs.empty()) { names = "."; coff.push_back(1); }
	ok = ok && CU_OK(cudaMalloc(&s->d_ctg, names.size() + 16)) && CU_OK(cudaMalloc(&s->d_ctg_off, coff.size() * sizeof(int))) && CU_OK(cudaMalloc(&s->d_nlines, 8)) &&
	     CU_OK(cudaMemcpyAsync(s->d_ctg, names.data(), names.size(), cudaMemcpyHostToDevice, c->st)) &&
	     CU_OK(cudaMemcpyAsync(s->d_ctg_off, coff.data(), coff.size() * sizeof(int), cudaMemcpyHostToDevice, c->st));
	int err = 0;
	ok = ok && CU_OK(cudaMemcpyAsync(&err, c->d_err_sites, sizeof(int), cudaMemcpyDeviceToHost, c->st)) && CU_OK(cudaStreamSynchronize(c->st));
	if (d_seg) cudaFree(d_seg);
	if (d_off) cudaFree(d_off);
	if (d_cnt) cudaFree(d_cnt);
	if (ok && (err & 512)) { set_err("corrupt BCF: record lengths do not chain to the end of the stream"); ok = false; }
	if (ok && (err & 1024)) { set_err("BCF record without INFO/_row or with fewer than two alleles (not the site side of a BGT database)"); ok = false; }
	if (ok) { // rows to the host: record windows of b200_view_text_ex are mapped to row ranges
		std::vector<SiteRec> h((size_t)s->n + 1);
		ok = s->n == 0 || CU_OK(cudaMemcpy(h.data(), s->d_sites, sizeof(SiteRec) * (size_t)s->n, cudaMemcpyDeviceToHost));
		s->h_rows.resize((size_t)s->n);
		for (int64_t i = 0; ok && i < s->n; ++i) { s->h_rows[(size_t)i] = h[(size_t)i].row; if (i && h[(size_t)i].row < h[(size_t)i - 1].row) s->rows_sorted = false; }
	}
	if (!ok) { b200_sites_destroy(s); return nullptr; }
	return s;
}

extern "C" int64_t b200_sites_n(const b200_sites_t *s) { return s ? s->n : -1; }
extern "C" int b200_sites_rows_sorted(const b200_sites_t *s) { return s && s->rows_sorted ? 1 : 0; }

// the records whose row lies in [row_beg, row_end): how a region shard (rows) maps to its records
extern "C" int b200_sites_rec_range(const b200_sites_t *s, int64_t row_beg, int64_t row_end, int64_t *rec_beg, int64_t *rec_end)
{
	if (!s || !rec_beg || !rec_end) { set_err("b200_sites_rec_range: null argument"); return -1; }
	if (!s->rows_sorted) { set_err_code(B200_E_UNORDERED_RECORDS, "b200_sites_rec_range: the records are not in row order"); return -1; }
	*rec_beg = std::lower_bound(s->h_rows.begin(), s->h_rows.end(), row_beg) - s->h_rows.begin();
	*rec_end = std::lower_bound(s->h_rows.begin(), s->h_rows.end(), row_end) - s->h_rows.begin();
	return 0;
}
extern "C" const char *b200_sites_header(const b200_sites_t *s, int64_t *len) { if (!s) return nullptr; if (len) *len = (int64_t)s->header.size(); return s->header.c_str(); }

// (site, row) table to the host: for callers that keep assembling records themselves
extern "C" int b200_sites_rows(const b200_sites_t *s, int64_t *rows, int32_t *pos)
{
	if (!s) return -1;
	cudaSetDevice(s->ctx->dev);
	std::vector<SiteRec> h((size_t)s->n + 1);
	if (s->n && !CU_OK(cudaMemcpy(h.data(), s->d_sites, sizeof(SiteRec) * (size_t)s->n, cudaMemcpyDeviceToHost))) return -1;
	for (int64_t i = 0; i < s->n; ++i) { if (rows) rows[i] = h[(size_t)i].row; if (pos) pos[i] = h[(size_t)i].pos; }
	return 0;
}

extern "C" int64_t b200_view_text_ex(b200_ctx_t *c, b200_sites_t *s, const b200_pbf_t *pb, const b200_query_t *q, unsigned flags,
                                     int64_t rec_beg, int64_t rec_end, const char *const *contig_names, int n_contigs, const char **text, int64_t *n_lines)
{
	if (!c || !s || !pb || !q || !text) { set_err("b200_view_text: null argument"); return -1; }
	if (s->ctx != c || pb->ctx != c || q->ctx != c) { set_err("b200_view_text: handles belong to another context"); return -1; }
	cudaSetDevice(c->dev);
	if (rec_end < 0 || rec_end > s->n) rec_end = s->n;
	if (rec_beg < 0) rec_beg = 0;
	if (rec_beg > rec_end) rec_beg = rec_end;
	const int64_t n_rec = rec_end - rec_beg;
	if (q->has_flt && q->prog.needs_host) { set_err_code(B200_E_FILTER_NEEDS_HOST, "b200_view_text: filters using ** are evaluated with the host libm; use b200_scan"); return -1; }
	bool with_counts = (flags & B200_VIEW_COUNTS) != 0;
	const bool with_gt = (flags & B200_VIEW_GENOTYPES) != 0;
	if (q->has_flt || q->G > 1) with_counts = true;                  // bgt.c:850: a filter or several groups imply AC/AN in the output
	// rows behind the records of this window
	int64_t row_lo = b200_pbf_row_beg(pb), row_hi = b200_pbf_row_end(pb);
	if (s->rows_sorted && n_rec > 0) { row_lo = s->h_rows[(size_t)rec_beg]; row_hi = s->h_rows[(size_t)rec_end - 1] + 1; }
	else if (!s->rows_sorted && (rec_beg != 0 || rec_end != s->n)) { set_err_code(B200_E_UNORDERED_RECORDS, "b200_view_text: the records are not in row order; only the whole file can be formatted at once"); return -1; }
	if (row_lo < b200_pbf_row_beg(pb) || row_hi > b200_pbf_row_end(pb)) { set_err("b200_view_text: rows [%lld,%lld) of the records are not resident", (long long)row_lo, (long long)row_hi); return -1; }
	const int64_t n_rows = row_hi - row_lo;
	if (contig_names && n_contigs > 0) { // the output header's contig dictionary (bgt.c:626-662) instead of the file's own
		bool same = (size_t)n_contigs == s->contigs.size();
		for (int i = 0; same && i < n_contigs; ++i) same = s->contigs[(size_t)i] == contig_names[i];
		if (!same) {
			std::string names; std::vector<int> coff(1, 0);
			for (int i = 0; i < n_contigs; ++i) { names += contig_names[i]; coff.push_back((int)names.size()); }
			cudaStreamSynchronize(c->st);
			if (s->d_ctg) cudaFree(s->d_ctg);
			if (s->d_ctg_off) cudaFree(s->d_ctg_off);
			s->d_ctg = nullptr; s->d_ctg_off = nullptr;
			if (!CU_OK(cudaMalloc(&s->d_ctg, names.size() + 16)) || !CU_OK(cudaMalloc(&s->d_ctg_off, coff.size() * sizeof(int))) ||
			    !CU_OK(cudaMemcpy(s->d_ctg, names.data(), names.size(), cudaMemcpyHostToDevice)) ||
			    !CU_OK(cudaMemcpy(s->d_ctg_off, coff.data(), coff.size() * sizeof(int), cudaMemcpyHostToDevice))) return -1;
			s->contigs.assign(contig_names, contig_names + n_contigs);
		}
	}
	// ---- the scan: per-row counts, verdicts and (for genotype columns) bit planes stay on the device
	const int stride = 3 + 3 * q->G;
	b200_scan_out_t so;
	memset(&so, 0, sizeof(so));
	unsigned sflags = B200_SCAN_DEVICE_OUT;
	if (with_counts && n_rows > 0) {
		if (!c->counts.reserve((size_t)n_rows * stride * sizeof(int32_t) + 16) || !c->pass.reserve((size_t)n_rows + 16)) return -1;
		so.counts = (int32_t*)c->counts.p; so.pass = (uint8_t*)c->pass.p; sflags |= B200_SCAN_COUNTS;
	}
	if (with_gt && n_rows > 0) {
		for (int p = 0; p < 2; ++p) { if (!c->hapbits[p].reserve((size_t)n_rows * q->words * sizeof(uint32_t) + 16)) return -1; so.hap_bits[p] = (uint32_t*)c->hapbits[p].p; }
		sflags |= B200_SCAN_HAP_BITS;
	}
	const bool scanned = (sflags & (B200_SCAN_COUNTS | B200_SCAN_HAP_BITS)) != 0;
	if (scanned && b200_scan(c, pb, q, row_lo, n_rows, sflags, &so) != n_rows) return -1;
	// ---- lines
	const size_t tb = view_scan_temp_bytes(n_rec);
	if (!s->len.reserve(8 * (size_t)(n_rec + 2)) || !s->off.reserve(8 * (size_t)(n_rec + 2)) || !s->temp.reserve(tb + 16)) return -1;
	ViewParams P;
	memset(&P, 0, sizeof(P));
	P.sites = s->d_sites + rec_beg; P.n_rec = n_rec; P.bcf = s->d_bcf; P.ctg_names = s->d_ctg; P.ctg_off = s->d_ctg_off; P.n_ctg = (int)(s->contigs.empty() ? 1 : s->contigs.size());
	P.counts = with_counts ? (const int32_t*)c->counts.p : nullptr; P.pass = q->has_flt ? (const uint8_t*)c->pass.p : nullptr; P.stride = stride; P.G = q->G; P.with_counts = with_counts ? 1 : 0;
	P.with_gt = with_gt ? 1 : 0; P.n_out = q->n_out; P.words = q->words; P.hap[0] = (const uint32_t*)c->hapbits[0].p; P.hap[1] = (const uint32_t*)c->hapbits[1].p;
	P.row_lo = row_lo; P.n_rows = n_rows; P.err = c->d_err_sites;
	unsigned long long total = 0, lines = 0;
	if (!CU_OK(cudaMemsetAsync(c->d_err_sites, 0, sizeof(int), c->st))) return -1;
	if (n_rec > 0) {
		bool ok = CU_OK(cudaMemsetAsync(s->len.p, 0, 8 * (size_t)(n_rec + 2), c->st)) && CU_OK(cudaMemsetAsync(s->d_nlines, 0, 8, c->st)) &&
		          CU_OK(launch_view_text(P, (unsigned long long*)s->len.p, (unsigned long long*)s->off.p, s->temp.p, tb, nullptr, nullptr, 0, c->st)) &&
		          CU_OK(cudaMemcpyAsync(&total, (unsigned long long*)s->off.p + n_rec, 8, cudaMemcpyDeviceToHost, c->st)) && CU_OK(cudaStreamSynchronize(c->st));
		c->launches += 2;
		if (!ok) return -1;
	}
	if (!s->text.reserve((size_t)total + 64)) return -1;
	if ((size_t)total + 1 > s->h_text_cap) {
		if (s->h_text) cudaFreeHost(s->h_text);
		s->h_text = nullptr; s->h_text_cap = 0;
		if (!CU_OK(cudaMallocHost((void**)&s->h_text, (size_t)total + 64))) return -1;
		s->h_text_cap = (size_t)total + 64;
	}
	int err = 0, serr = 0;
	unsigned long long vtot[4] = {0, 0, 0, 0};
	bool ok = (n_rec == 0 || CU_OK(launch_view_text(P, (unsigned long long*)s->len.p, (unsigned long long*)s->off.p, nullptr, 0, (char*)s->text.p, s->d_nlines, 1, c->st))) &&
	          (total == 0 || CU_OK(cudaMemcpyAsync(s->h_text, s->text.p, (size_t)total, cudaMemcpyDeviceToHost, c->st))) &&
	          CU_OK(cudaMemcpyAsync(&lines, s->d_nlines, 8, cudaMemcpyDeviceToHost, c->st)) && CU_OK(cudaMemcpyAsync(&err, c->d_err_sites, sizeof(int), cudaMemcpyDeviceToHost, c->st)) &&
	          CU_OK(cudaMemcpyAsync(&serr, c->d_err_scan, sizeof(int), cudaMemcpyDeviceToHost, c->st)) &&
	          CU_OK(cudaMemcpyAsync(vtot, c->d_acc, sizeof(vtot), cudaMemcpyDeviceToHost, c->st)) && CU_OK(cudaStreamSynchronize(c->st));
	++c->launches;
	if (!ok) return -1;
	if (scanned && serr) { set_err("device error flags 0x%x during scan", serr); return -1; }
	for (int i = 0; i < 4; ++i) c->last_totals[i] = scanned ? (int64_t)vtot[i] : 0;
	if (err & 2048) { set_err("a site record points at a row outside the scanned rows"); return -1; }
	if (err) { set_err("device error flags 0x%x while formatting", err); return -1; }
	s->h_text[total] = 0;
	*text = s->h_text;
	if (n_lines) *n_lines = (int64_t)lines;
	read_scan_timings(c);
	return (int64_t)total;
}

extern "C" int64_t b200_view_text(b200_ctx_t *c, b200_sites_t *s, const b200_pbf_t *pb, const b200_query_t *q, int with_counts,
                                  const char *const *contig_names, int n_contigs, const char **text, int64_t *n_lines)
{
	return b200_view_text_ex(c, s, pb, q, with_counts ? B200_VIEW_COUNTS : 0u, 0, -1, contig_names, n_contigs, text, n_lines);
}
