// pbwt_kernels.cuh -- device-side data structures and launch wrappers of the PBWT hot path (sm_100a).
//
// Data layout in HBM (all owned by b200_pbf_t, see api.cu):
//   img      : the .pbf byte image (or the byte range of the resident checkpoint blocks), 16-byte aligned
//              base, >= 64 bytes of zero padding behind it (TMA bulk copies are 16-byte granular).
//   rowoff   : uint64 [n_blk][BS+1], offset (relative to img) of the 'B' record of every row of every
//              resident block, entry BS (or rows_in_block) = end of the block's last row.  BS = 1<<shift.
//   n1       : uint32 [n_blk][BS][2], number of 1 bits of every (row, plane) = sum of its 1-run lengths.
//   tiles    : int2   [n_tiles], {first row in block, n_rows | big<<31}; blk_tile_beg int [n_blk+1].
//   rank0    : int32  [n_blk][2][m], rank of every column under the block's 'S' snapshot (inverse permutation).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "flt.h"

namespace b200 {

constexpr int WALK_NT   = 512;          // threads per CTA of the rank-walk kernel
constexpr int WALK_NW   = WALK_NT / 32;
constexpr int RAW_CAP   = 8192;         // RLE bytes staged per tile
constexpr int RAW_BYTES = RAW_CAP + 32; // + 16-byte alignment slack at both ends
constexpr int T_MAX     = 64;           // rows per tile
constexpr int COMP_K    = 32;           // rows per composite map (compose.cu); tiles never cross a multiple of it
constexpr int COMP_CAP  = 4096;         // pieces per composite map (<= RAW_BYTES: staged in the run-table buffers)
constexpr int COMP_DIR  = 2048;         // rank buckets of a composite map's directory (uint16 piece index per bucket + sentinel)
constexpr int COMP_DIR_STRIDE = COMP_DIR + 8;
constexpr int B200_MAX_GROUPS_K = 32;   // BGT_MAX_GROUPS, bgt.h:13

struct RowMeta { uint32_t off[2], len[2], n1[2]; };

// Two-sided composite maps (compose.cu, pairwalk.cu): in a full checkpoint block whose successor is resident too, the row
// groups of the second half get INVERSE maps (rank behind the group -> rank in front of it), so that a (column,row) pair is
// reached from the nearer of the two snapshots around it.  Returns the first group with an inverse map (n_grp: none).
__host__ __device__ inline int comp_first_inverse(int blk, int n_blk_res, int rows_in_blk, int BS, int n_grp, int two_sided)
{
	return (two_sided && blk + 1 < n_blk_res && rows_in_blk == BS) ? n_grp / 2 : n_grp;
}

struct WalkParams {
	const uint8_t  *img;
	const uint64_t *rowoff;
	const uint32_t *n1;
	const int2     *tiles;
	const int      *blk_tile_beg;   // [blocks] first tile of every block ...
	const int      *blk_tile_end;   // ... and one past its last tile (index.cu: the tiles of a block live in a fixed slot)
	const int32_t  *rank0;
	const int32_t  *track;     // tracked column ids [n_track] or nullptr (identity: every column)
	const uint8_t  *tgrp;      // 0-based group of every tracked column
	int32_t        *cnt_raw;   // [rows out][G][3] = #ALT, #missing, #other-ALT per group (zero-initialised, accumulated)
	uint32_t       *hap[2];    // [rows out][words] bit planes (EMIT only)
	const uint16_t *qrow;      // QUERY only: target row (within the block) of every tracked entry, ascending per block
	const uint32_t *comp_start; // QUERY only: composite maps (compose.cu) [blocks][groups][COMP_CAP], or nullptr
	const int32_t  *comp_delta;
	const int      *comp_n;     // [blocks][groups] pieces (padded to 4), 0 = not available
	const uint16_t *comp_dir;   // [blocks][groups][COMP_DIR_STRIDE]: last piece starting at or below rank (b << dir_shift), or nullptr
	int dir_shift, dir_n;       // bucket width 2^dir_shift; dir_n = directory entries incl. sentinel, padded to 8
	const int      *grp_tile_beg; // [blocks][groups+1] first tile of every row group
	const int      *blk_list;  // resident-block indices handled by this launch (nullptr: blk_first + blockIdx.y)
	const uint8_t  *blk_skip;  // nullptr, or per resident block 1 = not this kernel's (the block is on the split path): its CTAs exit
	const uint8_t  *blk_ok;    // nullptr, or per resident block 1 = this kernel's (QUERY: the block is on the split path)
	const int      *n_track_blk; // per-block number of tracked columns (nullptr: n_track)
	long long       track_stride; // > 0: track holds one list per resident block, this many entries apart
	uint8_t        *snap_img;  // CHAIN only: image to write 'S' snapshots into
	const uint64_t *blkoff;    // CHAIN only: offset of the 'S' record of every block
	int m, n_track, G, words, shift;
	int blk_first;             // resident-block index handled by blockIdx.y == 0
	int n_blk_chain;           // CHAIN only: blocks to run through
	long long blk_row0;        // absolute row of resident block 0
	long long row_lo, row_hi;  // absolute rows to produce output for
	int *err;
};

size_t walk_smem_bytes(int C, int G, bool query = false);
// C = tracked columns per thread (1,2,4,8); mode = WALK_COUNT / WALK_EMIT / WALK_CHAIN / WALK_QUERY (pbwt_kernels.cu)
enum { WALK_MODE_COUNT = 0, WALK_MODE_EMIT = 1, WALK_MODE_CHAIN = 2, WALK_MODE_QUERY = 3 };
cudaError_t launch_walk(const WalkParams &P, int C, int mode, int slices, int n_blk, cudaStream_t st);

// second phase of the split scan (pairwalk.cu): plane-0 bit of the (column,row) pairs that carry a plane-1 bit
struct PairParams {
	const uint8_t  *img;
	const uint64_t *rowoff;
	const uint32_t *n1;
	const int32_t  *rank0;       // [blocks][2][m] start ranks under every resident block's snapshot
	const int32_t  *qcol;        // pairs of every block, sorted by target row: column ...
	const uint16_t *qrow;        // ... and target row within the block
	const int      *qcount;      // [blocks] pairs per block
	long long       q_stride;    // entries between the lists of consecutive blocks
	const uint8_t  *tgrp;        // 0-based sample group of every column
	const uint32_t *comp_start;  // composite maps of the 32-row groups (compose.cu) [blocks][groups][COMP_CAP]
	const int32_t  *comp_delta;
	const int      *comp_n;      // [blocks][groups] pieces (padded to 4), 0 = not available; nullptr = no maps at all
	const uint16_t *comp_dir;    // [blocks][groups][COMP_DIR_STRIDE] bucket directories
	int dir_shift, dir_n;
	const int      *blk_list;    // resident-block indices handled by this launch (nullptr: blk_first + blockIdx.y)
	const uint8_t  *blk_ok;      // nullptr, or per resident block 1 = on the split path (device-side verdict of index.cu); others are skipped
	int blk_first;
	int32_t        *cnt_raw;     // [rows out][G][3] = #ALT, #missing, #other-ALT per group (accumulated)
	int m, G, shift;
	long long blk_row0, row_lo, row_hi;
	const int *rows_in_blk;
	int two_sided, n_blk_res;    // see comp_first_inverse
	int *err;
	unsigned long long *prof;    // diagnostics (BGT_B200_PROF): [0] CTA cycles in phase A, [1] in phase B, [2] CTAs, [3] composite groups crossed, [4] rows walked by warps
	int slice0;                  // this launch covers the slices from slice0 on (the extension launch of blocks flagged 2)
	const long long *ext_off;    // extension launch: qcol/qrow are the extension arrays, pair q of block b lives at ext_off[b] + q - ext_shift
	long long ext_shift;
};
constexpr int PAIR_SLICE_THREADS = 512;   // threads of a pair-walk / select CTA; a slice of a block's pair list = PAIR_SLICE_THREADS x C pairs
size_t pair_smem_bytes();
cudaError_t launch_pairwalk(const PairParams &P, int C, int max_pairs, int n_blk, cudaStream_t st);

// composite maps of row groups (compose.cu)
struct ComposeParams {
	const uint8_t  *img;
	const uint64_t *rowoff;
	const uint32_t *n1;
	const uint32_t *nrun;       // nullptr, or runs per row | bit of the first run << 31 (rowmeta_kernel): level 0 skips its counting pass
	int two_sided, n_blk_res;   // see comp_first_inverse: inverse maps for the second half of the groups of blocks with a resident successor
	const int      *rows_in_blk;
	const int      *blk_list;   // blocks handled by this launch; nullptr: blk_first + index
	const uint8_t  *blk_ok;     // nullptr, or per block 1 = build (the device-side "sparse" flag of index.cu)
	int blk_first;
	int m, shift;
	const long long *row_base; // per block: index of its first row in rowoff (with one extra end entry per block) / n1 / ...; nullptr: blk*(BS+1), blk*BS
	int n1_step;      // entries of n1 per row (2 for a .pbf image: both planes; 1 for the plane-1 view)
	int n_grp;        // row groups per block (slots per block in the output arrays)
	int cap;          // pieces per slot (<= COMP_CAP)
	int rle_off;      // offset of the plane's RLE inside a row record: 5 = plane 0 of a .pbf row, 9 = plane 1 of a plane-1 view row
	int n1_plane;     // which entry of n1[row][2] belongs to that plane
	int inverse;      // 1: inverse composite (coordinates behind the group -> in front of it)
	int retry;        // set by launch_compose: second launch, only the groups whose first attempt ran out of room
	uint16_t *comp_dir; // forward maps only: bucket directory per slot (COMP_DIR_STRIDE entries), or nullptr
	int dir_shift, dir_n;
	uint32_t *comp_start;
	int32_t  *comp_delta;
	int      *comp_n;
};
size_t compose_smem_bytes();
cudaError_t launch_compose(const ComposeParams &P, int n_blk, cudaStream_t st);

// plane-1 select (plane1.cu): per block, the (column, row) pairs that carry a plane-1 bit, in row order
struct SelectParams {
	const uint8_t  *p1img;        // plane-1 view: records 'B', l0 = 0, l1, bytes of the rows whose plane 1 is not empty
	const uint64_t *p1_rowoff;    // view rows of all blocks back to back, one extra end entry per block: index vbase[blk] + blk + v
	const uint32_t *p1_n1;        // ones of plane 1 per view row: index vbase[blk] + v
	const uint16_t *p1_realrow;   // row (within the block) of every view row: index vbase[blk] + v
	const long long *p1_vbase;    // [blocks+1] first view row of every block
	const int      *p1_rows_in_blk;
	const uint8_t  *img;          // the real image (for the block's plane-1 snapshot)
	const uint64_t *blkoff;
	const int      *blk_list;     // blocks handled by this launch; nullptr: blk_first + index
	const uint8_t  *blk_ok;       // nullptr, or per block 1 = sparse (select it)
	int blk_first;
	int m, shift, cap;
	const uint32_t *p1_prefix;    // plane-1 ones in front of every view row (one extra end entry per block): index vbase[blk] + blk + v
	const uint32_t *vcomp_start;  // inverse composites of the view rows (compose.cu, inverse = 1) [blocks][SELECT_GROUPS][SELECT_COMP_CAP], or nullptr
	const int32_t  *vcomp_delta;
	const int      *vcomp_n;      // [blocks][SELECT_GROUPS]
	const uint16_t *vcomp_dir;    // [blocks][SELECT_GROUPS][COMP_DIR_STRIDE] bucket directories
	int dir_shift, dir_n;
	int32_t  *qcol;               // out [blocks][cap]
	uint16_t *qrow;               // out [blocks][cap]
	int      *qcount;             // out [blocks]
	int      *err;
	int slice0, n_slices;         // this launch covers the slices [slice0, slice0 + n_slices) of every block's pair list (n_slices 0: all that cap holds)
	long long q_stride;           // entries between the lists of consecutive blocks in qcol/qrow (0: cap)
	const long long *ext_off;     // extension launch: qcol/qrow are the extension arrays, pair q of block b goes to ext_off[b] + q - ext_shift
	long long ext_shift;
};
// the plane-1 view of a block: up to SELECT_MAX_ROWS non-empty rows (every row of an ordinary block) in SELECT_MAX_BYTES of
// re-framed records; inverse composite maps of its 32-row groups with up to SELECT_COMP_CAP pieces
constexpr int SELECT_MAX_ROWS = 8192, SELECT_MAX_BYTES = 1 << 20;
constexpr int P1_SLOT_BYTES = SELECT_MAX_BYTES + 64;   // one block's slot in the plane-1 view image (index.cu)
constexpr int SELECT_GROUPS = SELECT_MAX_ROWS / COMP_K, SELECT_COMP_CAP = 2048;
cudaError_t launch_plane1_select(const SelectParams &P, int n_blk, cudaStream_t st);

// row index of a resident PBF, built on the device (index.cu)
struct IndexParams {
	const uint8_t  *img;
	const uint64_t *blkoff;       // [blocks] offset of the 'S' record of every resident block, relative to img
	const uint64_t *blkend;       // [blocks] end of the block's records (next 'S' record or the 'I' record)
	int            *rows_in_blk;  // [blocks] rows per block; set to 0 by the index kernel if the block's records do not parse
	int m, shift, blk_first;
	uint64_t *rowoff;             // out [blocks][BS+1]
	uint64_t *scratch;            // [blocks][IX_SCRATCH_LANES][BS+1] per-lane pieces of the chain, or nullptr (single-lane chase)
	int2     *tiles;              // out [blocks][BS]
	int      *blk_tile_beg, *blk_tile_end, *grp_tile_beg;  // out [blocks], [blocks], [blocks][groups+1]
	int      *err;
	int      *fallbacks;          // diagnostics: blocks whose pieces did not join up (chased again by one lane)
};
constexpr int IX_SCRATCH_LANES = 12;
cudaError_t launch_index(const IndexParams &P, int n_blk, cudaStream_t st);
cudaError_t launch_plan_tiles(const IndexParams &P, int n_blk, cudaStream_t st);

struct P1ViewParams {
	const uint8_t  *img;
	const uint64_t *rowoff;
	const uint32_t *n1;
	const int      *rows_in_blk;
	int m, shift, blk_first, p1_cap;
	int p1_base;                  // blocks with more plane-1 ones than this (and at most p1_cap) are flagged 2 instead of 1 (see blk_sparse)
	uint32_t  *blk_ones;          // out [blocks] plane-1 ones (= (column,row) pairs) of the block
	uint8_t   *p1img;             // out [blocks][P1_SLOT_BYTES]
	uint64_t  *p1_rowoff;         // out [blocks][SELECT_MAX_ROWS+1]
	uint32_t  *p1_n1;             // out [blocks][SELECT_MAX_ROWS]
	uint32_t  *p1_prefix;         // out [blocks][SELECT_MAX_ROWS+1] ones of plane 1 in front of every view row
	uint16_t  *p1_realrow;        // out [blocks][SELECT_MAX_ROWS]
	int       *p1_rows_in_blk;    // out [blocks]
	long long *p1_vbase;          // out [blocks] = blk * SELECT_MAX_ROWS
	uint8_t   *blk_sparse;        // out [blocks] != 0: the split scan applies to the block.  1: its pairs fit the launches that are sized
	                              // without knowing the data (p1_base: the load pipeline's); 2: they need the extension launches (slices behind
	                              // p1_base, queued once the host has seen blk_ones)
};
cudaError_t launch_p1view(const P1ViewParams &P, int n_blk, cudaStream_t st);

// per-group plane-0 marginals (marginal.cu): n0g[row][g] = ones of the plane-0 row among the columns of group g
struct MarginalParams {
	const uint8_t  *img;
	const uint64_t *rowoff;
	const uint32_t *n1;
	const uint64_t *blkoff;
	const int      *rows_in_blk;
	const int      *blk_list;  // nullptr: blk_first + launch index
	const uint8_t  *blk_ok;    // nullptr, or per resident block 1 = on the split path; the CTAs of other blocks exit
	int blk_first;
	const uint8_t  *tgrp;      // 0-based group per column (full-cohort queries: tracked entry == column)
	int32_t        *n0g;       // out [rows out][n_vec]
	int m, shift, n_vec;
	long long blk_row0, row_lo, row_hi;
	// segmented run (n_seg > 1): the blocks' vectors are first pushed through the composite maps of the row groups and stored
	// in front of every segment of seg_groups groups (vseg [launch block][vector][segment][marginal_seg_words(m)]); every
	// segment is then walked by its own CTA.  seg_ok [launch block][vector] = 0: a composite was missing, one CTA takes the block
	const uint32_t *comp_start; const int32_t *comp_delta; const int *comp_n;
	int two_sided, n_blk_res;   // see comp_first_inverse: the maps of those groups are gathered through instead of scattered through
	int n_grp, seg_groups, n_seg;
	uint32_t *vseg; uint8_t *seg_ok;
	// slots of one (block, vector) in vseg and the distance between the slots of two consecutive segments.  n_seg, 1: the vectors in
	// front of the segments only.  n_grp + 1, seg_groups: `dense` vectors, one in front of every 32-row group and one behind the
	// last full group (what margpiece.cu works from; the row loop of marginal.cu reads every seg_groups-th of them)
	int seg_slots, seg_slot_step;
	// margpiece.cu
	const uint32_t *nrun;      // runs of plane 0 per row | first bit << 31 (rowmeta_kernel)
	uint8_t *blk_fail;         // [launch block][vector] raised by margpiece.cu: redo this block with the row loop
	uint4 *vrec;               // look-up records of the dense vectors: [launch block][vector][slot][marginal_rec_words64(m)] {64 bits, members in front}
	uint2 *retry; int *retry_n; int retry_cap;   // chains the small configuration of margpiece.cu gave up on
	const uint8_t *blk_only;   // the row loop: nullptr, or [launch block][vector] 1 = run (everything else exits)
};
size_t marginal_smem_bytes(int m);
size_t marginal_seg_words(int m);
__host__ __device__ inline size_t marginal_seg_words_dev(int m) { return (size_t)(((m + 31) / 32 + 4 + 3) & ~3); }
cudaError_t launch_marginal(const MarginalParams &P, int n_blk, cudaStream_t st);
size_t marginal_rec_words64(int m);
__host__ __device__ inline size_t marginal_rec_words64_dev(int m) { return (size_t)((m + 63) / 64) + 1; }
cudaError_t launch_marginal_dense_seed(const MarginalParams &P, int n_blk, cudaStream_t st);
cudaError_t launch_marginal_seed(const MarginalParams &P, int n_blk, cudaStream_t st);
cudaError_t launch_marginal_rows(const MarginalParams &P, int n_blk, cudaStream_t st);
cudaError_t launch_marginal_pieces(const MarginalParams &P, int n_blk, cudaStream_t st);

// split scan: rows of blocks flagged in blk_split take #ALT of group g from the plane-0 marginal -- n0g[row][g] for the
// first n_vec groups, the rest of n1[row][0] for the last group -- minus the group's other-ALT count
struct FinalizeSplit { const uint8_t *blk_split; const uint32_t *n1; const int32_t *n0g; int n_vec; long long row_lo, blk_row0; int shift; };

cudaError_t launch_rowmeta(const uint8_t *img, const uint64_t *rowoff, int n_blk, int shift, long long n_rows_total_in_blocks,
                           const int *rows_in_blk, uint32_t m, uint32_t *n1, uint32_t *nrun0, unsigned long long *bad, cudaStream_t st);
cudaError_t launch_invert_snapshots(const uint8_t *img, const uint64_t *blkoff, int n_blk, int m, int32_t *rank0, int *err, cudaStream_t st);
cudaError_t launch_finalize(const int32_t *cnt_raw, long long n_rows, int G, const int32_t *gsize, const flt_prog_t *prog, int use_flt,
                            int32_t *counts, uint8_t *pass, unsigned long long *totals, const FinalizeSplit &sp, cudaStream_t st);
cudaError_t launch_unpack_bits(const uint32_t *bits, long long n_rows, int words, int n_track, uint8_t *bytes, cudaStream_t st);

// PBWT encoder (encode.cu)
struct EncodeParams {
	const uint32_t *in_bits;       // [n_rows][2][words] the batch's rows as bit planes in column order
	int m, words, shift, n_rows;
	long long row0;                // absolute index of the batch's first row
	int32_t  *rank;                // [2][m] rank of every column per plane: encoder state, carried from batch to batch
	uint32_t *bitvec;              // [3][2][words] rotating scatter targets (all zero before the first row)
	int32_t  *snap;                // out: 'S' snapshots of the checkpoints inside the batch, [k][2][m]
	uint8_t  *out[2];              // out: run-length bytes of the rows, one stream per plane
	uint32_t *row_len;             // out [n_rows][2]: bytes per (row, plane)
	const unsigned long long *out_pos0;   // [2] stream offsets at the start of the batch ...
	unsigned long long *out_pos1;         // ... and at its end
};
size_t encode_smem_bytes(int m);
cudaError_t launch_encode(const EncodeParams &P, int sm_count, cudaStream_t st);
cudaError_t launch_pack_rows(const uint8_t *a0, const uint8_t *a1, long long n_rows, int m, uint32_t *bits, cudaStream_t st);

// BGZF inflate (inflate.cu)
struct InflateParams {
	const uint8_t  *in;         // the compressed file image
	const uint64_t *blk_coff;   // [blocks] offset of the block's DEFLATE stream in `in`
	const uint32_t *blk_csize;  // [blocks] bytes of that stream
	const uint32_t *blk_usize;  // [blocks] uncompressed size (ISIZE, bgzf.c:251-257)
	const uint64_t *blk_uoff;   // [blocks] offset of the block's bytes in `out`
	uint8_t *out;
	uint32_t max_csize;
	int *err;
};
cudaError_t launch_bgzf_inflate(const InflateParams &P, int n_blk, cudaStream_t st);

// site side of `bgt view` (sites.cu)
struct SiteRec {
	int32_t rid, pos, rlen, n_allele;
	int32_t ref_len, alt_len;
	unsigned long long ref_off, alt_off;   // where REF and the first ALT lie in the inflated BCF stream
	long long row;                         // INFO/_row (bgt.c:279-286); -1 = the record did not parse
};
struct ViewParams {
	const SiteRec *sites; long long n_rec;
	const uint8_t *bcf;
	const char *ctg_names; const int *ctg_off; int n_ctg;   // contig names back to back, ctg_off[n_ctg+1]
	const int32_t *counts; const uint8_t *pass;              // the scan's per-row results (rows row_lo .. row_lo+n_rows)
	int stride, G, with_counts;
	int with_gt, n_out, words;                               // genotype columns: samples, words per plane row of hap[]
	const uint32_t *hap[2];                                  // the scan's bit planes [n_rows][words]
	long long row_lo, n_rows;
	int *err;
};
cudaError_t launch_bcf_chase(const uint8_t *bcf, unsigned long long bcf_len, const unsigned long long *seg_pos, int n_seg, int seg_len, long long n_rec,
                             unsigned long long *rec_off, unsigned long long *counted, int *err, cudaStream_t st);
cudaError_t launch_bcf_parse(const uint8_t *bcf, unsigned long long bcf_len, const unsigned long long *rec_off, long long n_rec, int row_key, SiteRec *sites,
                             int *err, cudaStream_t st);
size_t view_scan_temp_bytes(long long n);
cudaError_t launch_view_text(const ViewParams &P, unsigned long long *len, unsigned long long *off, void *temp, size_t temp_bytes, char *text,
                             unsigned long long *n_lines, int phase, cudaStream_t st);

// synthetic cohort generator (synth.cu)
struct SynthCfg { uint32_t m; long long n_rows; int shift; uint64_t seed; int r_max; int p1_one_in; int p1_max_iv; int p1_max_len; };
cudaError_t launch_synth_lengths(const SynthCfg &c, uint32_t *len2 /*[n_rows][2]*/, cudaStream_t st);
cudaError_t launch_synth_write(const SynthCfg &c, const uint64_t *rowoff_flat /*[n_rows] absolute*/, uint8_t *img, cudaStream_t st);

} // namespace b200
