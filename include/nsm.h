/*
 * nsm.h — C ABI of the B200-native cross-cohort comparison path (libnsm_b200.so).
 *
 * The reference (BIH-CEI/napkon-string-matching) is pure Python and has no FFI; the seam this
 * library sits behind is the body of
 *     ComparableData.gen_comparable      napkon_string_matching/types/comparable_data.py:223-243
 * i.e. "for every (left item, right item) of the cross product (:191): compare_terms(...)
 * (:248-265) with score_func = intersection_vs_union | fuzzy_match
 * (compare/score_functions.py:6-27), keep MatchScore >= score_threshold (:243)".
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add at that point.
 *
 * Conventions
 *  - Every pointer inside nsm_sets_t / nsm_strings_t / nsm_job_t is a DEVICE pointer on the
 *    current CUDA device.  The library never allocates, frees or retains caller-visible memory;
 *    it keeps no state besides a thread-local error string.
 *  - Calls enqueue work on `stream` (a cudaStream_t passed as void*) and return without
 *    synchronising.  out_count / out_flags / out_stats are zeroed on the stream first; read
 *    them after synchronising the stream.
 *  - Return value: NSM_OK or an NSM_ERR_* code; nsm_last_error() gives the text.  Nothing throws.
 *  - There is no CPU implementation behind these entry points.
 */
#ifndef NSM_H_
#define NSM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSM_VERSION 100 /* 0.1.0 */

/* return codes */
#define NSM_OK 0
#define NSM_ERR_BAD_ARG 1
#define NSM_ERR_CUDA 2
#define NSM_ERR_UNSUPPORTED 3

/* bits of *out_flags */
#define NSM_FLAG_OVERFLOW 1u   /* more pairs kept than out_capacity (count is still exact) */
#define NSM_FLAG_ZERO_UNION 2u /* an evaluated level pair had two empty sets: the reference raises
                                  ZeroDivisionError (score_functions.py:13); pair not emitted */
#define NSM_FLAG_EMPTY_ITEM 4u /* an item without levels met one with levels: the reference raises
                                  IndexError (comparable_data.py:262); pair not emitted */

/* cat_mode: categories_matching, comparable_data.py:464-476 */
#define NSM_CAT_OFF 0
#define NSM_CAT_LIST_LIST 1 /* keep if masks intersect or both are empty */
#define NSM_CAT_MEMBER 2    /* keep if masks intersect (x in set(y), x == y) */

/* One kept pair: positions in the left / right cohort and the float64 MatchScore. */
typedef struct nsm_pair {
    uint32_t left;
    uint32_t right;
    double score;
} nsm_pair_t;

/* out_mode: how kept pairs are written */
#define NSM_OUT_PAIRS 0   /* out_pairs is nsm_pair_t[out_capacity]: 16 bytes per kept pair (default) */
#define NSM_OUT_PACKETS 1 /* out_pairs is nsm_packet_t[out_capacity] (nsm_jaccard_allpairs only):
                             ~10.4 bytes per kept pair for results that are bound by the
                             device->host link.  The kept pairs of one packet lie in one
                             512 x 128 block of the cross product. */
#define NSM_OUT_CODED 2   /* out_pairs is nsm_cpacket_t[out_capacity] (nsm_jaccard_allpairs only):
                             4.3 bytes per kept pair.  The float64 score of a pair is replaced by a
                             16-bit code: its slot in the score dictionary out_dict, an open-addressing
                             hash table over the score's bit pattern that the kernel fills first come,
                             first served (dense results hold few distinct scores: 99 % of the 1.4e8
                             kept pairs of a 50k x 50k TokenIds comparison share < 65536 values).  A
                             pair whose score finds no slot goes to out_exc as a plain nsm_pair_t. */
#define NSM_PACKET_RECORDS 48
#define NSM_CPACKET_RECORDS 60
#define NSM_DICT_SLOTS 65536u
#define NSM_DICT_FREE 0xffffffffffffffffull /* a free dictionary slot (not the bits of any score) */

/* Up to 48 kept pairs of the block (left0 .. left0+511) x (right0 .. right0+127):
 * pair i (i < count) is (left0 + (local[i] >> 7), right0 + (local[i] & 127), score[i]). */
typedef struct nsm_packet {
    uint32_t left0;
    uint32_t right0;
    uint32_t count;
    uint32_t reserved_;
    double score[NSM_PACKET_RECORDS];
    uint16_t local[NSM_PACKET_RECORDS];
} nsm_packet_t; /* 496 bytes */

/* NSM_OUT_CODED: pair i (i < count) is (left0 + ((rec[i] & 0xffff) >> 7), right0 + (rec[i] & 127),
 * the float64 whose bit pattern is out_dict[rec[i] >> 16]). */
typedef struct nsm_cpacket {
    uint32_t left0;
    uint32_t right0;
    uint32_t count;
    uint32_t reserved_;
    uint32_t rec[NSM_CPACKET_RECORDS];
} nsm_cpacket_t; /* 256 bytes */

/* One cohort side for intersection_vs_union: CSR items -> levels -> sorted unique token ids
 * (layout and meaning: napkon_string_matching/gpu/pack.py).  Levels are what
 * ComparableData.gen_comp_value returns per item (comparable_data.py:283-285). */
typedef struct nsm_sets {
    const uint32_t *item_level_off; /* [n_items + 1] */
    const uint32_t *level_tok_off;  /* [n_levels + 1] */
    const uint32_t *tok;            /* [level_tok_off[n_levels]] ids ranked by falling frequency */
    const uint8_t *tok_entry;       /* [same] first level of the item that holds the id (see nested) */
    const uint64_t *level_head;     /* [n_levels] exact bitset of the level's ids 0..63 */
    const uint64_t *level_tail;     /* [n_levels] signature of its ids >= 64 (exact iff exact_bits) */
    const uint64_t *level_tail2;    /* [n_levels] second, independent signature of the ids >= 64 */
    const uint32_t *level_info;     /* [n_levels] size | min(n_tail - popc(tail), 255) << 16 */
    const uint64_t *item_any;       /* [n_items][2] OR of (head, tail) over the levels compare_terms uses */
    const uint32_t *item_k;         /* [n_items] number of levels */
    const uint64_t *slot_ht;        /* [n_slots][slot_stride][2] (head, tail) of level min(t, K-1), slot t-1 */
    const uint32_t *slot_info;      /* [n_slots][slot_stride] level_info of the same level */
    uint32_t n_items;
    uint32_t n_levels;
    uint32_t max_levels; /* max levels of any item on this side */
    uint32_t n_slots;    /* clamp(max_levels - 1, 1, 10) */
    uint32_t exact_bits; /* 1: vocabulary <= 128 ids, tail bit == id - 64, no token merge needed */
    uint32_t slot_stride; /* items per slot row: n_items rounded up to 128, zero-filled */
    uint32_t nested;      /* 1: in every item level j is a subset of level j+1 (what gen_comp_value
                             produces), so one intersection of the deepest levels with tok_entry
                             yields the intersections of all level pairs */
    uint32_t reserved_;
} nsm_sets_t;

/* One cohort side for fuzzy_match: per level the processed string QRatio sees
 * (join_sorted + default_process, score_functions.py:16-27), one alphabet code per code point. */
typedef struct nsm_strings {
    const uint32_t *item_level_off; /* [n_items + 1] */
    const uint32_t *level_chr_off;  /* [n_levels] start of the level's string in chr, multiple of 8 */
    const uint32_t *level_len;      /* [n_levels] length in code points */
    const uint8_t *chr;             /* codes < n_alphabet; every string padded to 8 bytes */
    const uint32_t *level_hist;     /* [n_levels][8]: 32 byte counters (saturating at 255) of the level
                                       string's codes, bucket = code & 31 (byte b of word q = bucket
                                       4 q + b): the sound distance bound of the flat kernel */
    uint32_t n_items;
    uint32_t n_levels;
    uint32_t max_levels;
    uint32_t max_len;    /* longest level string on this side */
    uint32_t n_alphabet; /* codes in use, shared by both sides, <= 255 (code points that occur on one
                            side only share one code per side: they can never match) */
    uint32_t reserved_;
    /* Items are ordered by the length of their longest level string, so also by the 64-bit words it
     * needs: items [class_end[w-1], class_end[w]) need w+1 words; items from class_end[7] on hold a
     * level string of more than 512 characters.  The kernels run every class of the right side with
     * its own width; pairs with a long item go through the swapped / warp-cooperative passes. */
    uint32_t class_end[8];
} nsm_strings_t;

/* What to compare and where the kept pairs go. */
typedef struct nsm_job {
    uint32_t l_row_begin; /* left items [l_row_begin, l_row_end) x all right items: the row block */
    uint32_t l_row_end;   /*   one GPU owns (multi-GPU partitioning is by left row blocks) */
    uint32_t flat;        /* 0: compare_terms over levels (weights 2^-i, index from 1);
                             1: items have one level, score = score_func(level0, level0) */
    uint32_t cat_mode;    /* NSM_CAT_* ; masks below may be NULL when NSM_CAT_OFF */
    double threshold;     /* keep score >= threshold (float64 compare, comparable_data.py:243) */
    const uint64_t *l_cat; /* [left.n_items] category bit masks */
    const uint64_t *r_cat; /* [right.n_items] */
    void *out_pairs;       /* nsm_pair_t / nsm_packet_t / nsm_cpacket_t [out_capacity] (out_mode),
                              16-byte aligned, filled densely in no particular order */
    uint64_t out_capacity;
    uint64_t *out_count; /* number of kept pairs (packet modes: of packets), also beyond capacity */
    uint32_t *out_flags; /* NSM_FLAG_* */
    uint64_t *out_stats; /* [NSM_N_STATS] counters; may be NULL with NSM_OUT_PAIRS */
    uint32_t out_mode;   /* NSM_OUT_* */
    uint32_t reserved_;
    /* NSM_OUT_CODED only (ignored otherwise) */
    uint64_t *out_dict;        /* [NSM_DICT_SLOTS] score dictionary.  NOT reset by the library: the
                                  caller fills it with NSM_DICT_FREE before the first call of a
                                  result and may share it between the row blocks of that result */
    nsm_pair_t *out_exc;       /* [out_exc_capacity] pairs whose score found no dictionary slot */
    uint64_t out_exc_capacity;
    uint64_t *out_exc_count;   /* zeroed by the library; counts beyond the capacity too
                                  (NSM_FLAG_OVERFLOW is then set) */
} nsm_job_t;

#define NSM_STAT_CANDIDATES 0   /* item pairs that reached exact float64 scoring */
#define NSM_STAT_LEVEL_EVALS 1  /* score_func evaluations done exactly (popcount / merge / LCS) */
#define NSM_STAT_LEVEL_MERGES 2 /* of those, ones that needed a token merge */
#define NSM_STAT_BOUND_PAIRS 3  /* item pairs that reached the per-level bound (shared a signature bit) */
#define NSM_STAT_KEPT 4         /* kept pairs (== *out_count with NSM_OUT_PAIRS) */
#define NSM_N_STATS 5

int nsm_version(void);
const char *nsm_last_error(void);
/* Number of kernels the last nsm_*_allpairs / nsm_microbench call of this thread launched
 * (the fuzzy path launches one kernel per word-count class of the right side). */
int nsm_last_launch_count(void);

/* Replaces the pair loop of gen_comparable (comparable_data.py:223-243) for
 * score_func == "intersection_vs_union" (score_functions.py:6-13). */
int nsm_jaccard_allpairs(const nsm_sets_t *left, const nsm_sets_t *right, const nsm_job_t *job,
                         void *stream);

/* Same for score_func == "fuzzy_match" (score_functions.py:20-27; rapidfuzz QRatio/100 as the
 * normalised Indel similarity, bit-parallel LCS). */
int nsm_qratio_allpairs(const nsm_strings_t *left, const nsm_strings_t *right, const nsm_job_t *job,
                        void *stream);

/* Copies `bytes` (<= 256, a multiple of 8) of device memory to PINNED host memory with a one-warp
 * kernel instead of the copy engine: the counters of a finished call (out_count, out_flags,
 * out_stats) reach the host while a large record copy of an earlier call still occupies the
 * device->host copy engine.  host_dst: page-locked host memory (device-accessible under UVA). */
int nsm_publish(const void *dev_src, void *host_dst, uint32_t bytes, void *stream);

/* Host-side decoders of the compact record formats (plain CPU loops, no GPU involved: the host
 * half of NSM_OUT_PACKETS / NSM_OUT_CODED for any binding).  `out` (host memory) must hold the sum
 * of the packets' counts; `left_perm` / `right_perm` (or NULL) map the stored positions in the
 * records to the caller's item indices.  Return the number of records written. */
uint64_t nsm_decode_packets(const nsm_packet_t *packets, uint64_t n_packets, const uint32_t *left_perm,
                            const uint32_t *right_perm, nsm_pair_t *out);
uint64_t nsm_decode_cpackets(const nsm_cpacket_t *packets, uint64_t n_packets, const uint64_t *dict,
                             const uint32_t *left_perm, const uint32_t *right_perm, nsm_pair_t *out);

/* Host-side: sorts `n` records by (left, right) — the row-major order of the cross product, the
 * order of the reference's result frame (comparable_data.py:191) — into `sorted` (host memory, n
 * records, must not overlap `records`): counting sort by left, then each left row by right.
 * Every left index must be < n_left.  Returns 0, or NSM_ERR_BAD_ARG. */
int nsm_sort_pairs(const nsm_pair_t *records, uint64_t n, uint32_t n_left, nsm_pair_t *sorted);

/* Marks every slot of a score dictionary (NSM_OUT_CODED, uint64[NSM_DICT_SLOTS]) free. */
int nsm_dict_reset(uint64_t *dict, void *stream);

/* ---- device-side token packing ---------------------------------------------
 * Builds every array of nsm_sets_t on the GPU from the dictionary codes of the tokens, i.e. what
 * ComparableData.gen_comp_value (comparable_data.py:283-285) yields after the host has tokenised
 * the items and mapped token strings to integer codes (exact string identity, Q3).  The result is
 * bit-identical to the host packer (napkon_string_matching/gpu/pack.py:finish_sets). */
#define NSM_RAW_SUFFIX_PARTS 0 /* group g of an item is one part of its value; level j is the id set
                                  of the item's last j+1 parts (items[-i:], comparable_data.py:284) */
#define NSM_RAW_LEVELS 1       /* group g of an item is its level j, given explicitly */

#define NSM_PACK_MAX_ITEM_IDS 1024u /* ids one item may hold over all its groups */

/* bits of *flags written by the pack entry points */
#define NSM_PACK_FLAG_TOO_LARGE 1u  /* an item holds more than NSM_PACK_MAX_ITEM_IDS ids */
#define NSM_PACK_FLAG_NOT_NESTED 2u /* some level is not a subset of the next one (never with parts) */
#define NSM_PACK_FLAG_BAD_ID 4u     /* an id >= n_vocab */

typedef struct nsm_raw_sets {
    const uint32_t *item_grp_off; /* [n_items + 1] groups (parts or levels) of item i */
    const uint32_t *grp_id_off;   /* [n_groups + 1] ids of group g */
    const uint32_t *ids;          /* [n_ids] dictionary codes < n_vocab, any order, duplicates allowed */
    const uint32_t *rank;         /* [n_vocab] renumbering applied to every id (frequency rank), or NULL */
    uint32_t n_items;
    uint32_t n_groups;
    uint32_t n_ids;
    uint32_t n_vocab;
    uint32_t mode; /* NSM_RAW_* */
    uint32_t reserved_;
} nsm_raw_sets_t;

/* counts[id] += occurrences of id in raw->ids (raw->rank is ignored).  counts: device, [n_vocab],
 * zeroed by the caller before the first side is counted. */
int nsm_pack_count_ids(const nsm_raw_sets_t *raw, uint32_t *counts, void *stream);

/* Bytes of device scratch nsm_pack_sets_measure needs for n_items items. */
uint64_t nsm_pack_scratch_bytes(uint32_t n_items);

/* Pass 1: item_tok_off[i] (device, [n_items + 1]) = number of token rows (sum of the level sizes
 * after sort + unique) of all items before i; totals[0] = their total, totals[1] = NSM_PACK_FLAG_*
 * (device uint64[2]).  The caller reads totals, allocates the nsm_sets_t arrays and calls fill. */
int nsm_pack_sets_measure(const nsm_raw_sets_t *raw, uint32_t *item_tok_off, uint64_t *totals,
                          void *scratch, uint64_t scratch_bytes, void *stream);

/* Pass 2: fills every array `out` points to (device memory of the sizes nsm_sets_t documents, with
 * n_levels == raw->n_groups, tok / tok_entry of totals[0] rows; written through the const
 * pointers).  The caller sets out->n_items, n_levels, max_levels, n_slots, slot_stride and
 * exact_bits; out->nested is not touched (it is !(flags & NSM_PACK_FLAG_NOT_NESTED) &&
 * max_levels <= 255).  flags: device uint32, OR of NSM_PACK_FLAG_*. */
int nsm_pack_sets_fill(const nsm_raw_sets_t *raw, const uint32_t *item_tok_off, const nsm_sets_t *out,
                       uint32_t *flags, void *stream);

/* ---- device-side string packing (second half) -------------------------------
 * Applies the default processor of fuzz.QRatio (score_functions.py:27; rapidfuzz 2.1.x
 * utils.default_process: every non-alphanumeric code point becomes a blank, the ends are trimmed,
 * the rest is lower-cased) to the level strings join_sorted produced (score_functions.py:16-17)
 * and writes the `chr` and `level_hist` arrays of nsm_strings_t, bit-identical to the host packer
 * (napkon_string_matching/gpu/pack.py:pack_strings).  The host keeps what is per distinct code
 * point (its processed form, Python's Unicode tables) and per level (ordering, offsets); the GPU
 * does what is per character. */
#define NSM_STR_SYM_NONE 0xffffu /* cp_sym entry of a code point the host did not map */
#define NSM_STR_FLAG_UNMAPPED 1u /* a code point >= table_len or mapped to NSM_STR_SYM_NONE */

typedef struct nsm_raw_strings {
    const uint32_t *level_off; /* [n_levels + 1] start of level g's code points in cps */
    const uint32_t *cps;       /* [n_cps] UTF-32 code points of the unprocessed level strings */
    const uint16_t *cp_sym;    /* [table_len] code point -> symbol (index of its processed form) */
    uint32_t n_levels;
    uint32_t n_cps;
    uint32_t table_len;
    uint32_t blank_sym; /* the symbol of ' ': what the trim removes */
} nsm_raw_strings_t;

/* Pass 1: first[g] = index of level g's first character that survives the trim, len[g] = length of
 * the processed string (0 when nothing survives); both device uint32[n_levels].  flags: device
 * uint32, OR of NSM_STR_FLAG_*. */
int nsm_pack_strings_measure(const nsm_raw_strings_t *raw, uint32_t *first, uint32_t *len,
                             uint32_t *flags, void *stream);

/* Pass 2: for every STORED level s (the host ordered the items by the length of their longest
 * level, pack.py:pack_strings): the codes sym_code[cp_sym[c]] of the characters
 * first[src_level[s]] .. + level_len[s] of raw level src_level[s] go to chr + level_chr_off[s],
 * zero-padded to a multiple of 8 bytes, and their 32 saturating byte counters to
 * level_hist[s][8].  src_level, level_chr_off, level_len: device uint32[n_stored];
 * sym_code: device uint8[n_sym]. */
int nsm_pack_strings_fill(const nsm_raw_strings_t *raw, const uint32_t *first, const uint32_t *src_level,
                          const uint32_t *level_chr_off, const uint32_t *level_len,
                          const uint8_t *sym_code, uint32_t n_stored, uint8_t *chr,
                          uint32_t *level_hist, void *stream);

/* Integer-pipe micro-benchmarks used as roofline denominators: every thread of
 * a blocks x threads grid runs `iters` rounds of 8 independent chains x 4 dependent ops of
 * `kind` (0: LOP3, 1: IADD3, 2: POPC, 3: 64-bit add/sub/and/or LCS step).  *ops_per_thread
 * receives the number of 32-bit integer ops one thread executed.  Time it with CUDA events. */
int nsm_microbench(int kind, uint32_t blocks, uint32_t threads, uint32_t iters, uint32_t *sink,
                   uint64_t *ops_per_thread, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NSM_H_ */
