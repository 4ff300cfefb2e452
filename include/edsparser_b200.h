/*
 * edsparser_b200 — C ABI of the B200-native EDSParser hot path (libedsparser_b200.so).
 *
 * Plain C, POD structs, status codes, no exceptions and no torch types across the boundary.
 * The reference (draessld/EDSParser) has no FFI of its own: its boundary for this path is the
 * public C++ API in src/cpp/lib/transforms/ + src/cpp/lib/formats/. Each entry point below names
 * the reference interface it stands behind; edsparser_b200/host/ holds the C++17 wrappers that
 * keep those signatures and call this ABI.
 *
 * There is NO CPU fallback: every transform entry point needs a CUDA device and returns
 * EDS_ERR_CUDA when none is usable.
 */
#ifndef EDSPARSER_B200_H
#define EDSPARSER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes map 1:1 onto the exception classes the reference throws on this path
 * (SURVEY.md §8b): the C++ wrappers rethrow them with eds_last_error() as what(). */
typedef enum eds_status {
    EDS_OK = 0,
    EDS_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument (eds_transforms.cpp:322-324,388-390,395-397) */
    EDS_ERR_RUNTIME = 2,          /* std::runtime_error   (eds.cpp:80-82,123-131,278-349,1513-1519)   */
    EDS_ERR_OUT_OF_RANGE = 3,     /* std::out_of_range    (std::stoi overflow in eds.cpp:302,322)      */
    EDS_ERR_CUDA = 4,             /* CUDA runtime failure or no device (no CPU fallback exists)        */
    EDS_ERR_BAD_MSA = 5,          /* input outside the reference's well-defined MSA domain (C.2)       */
    EDS_ERR_BUDGET = 6,           /* output larger than the caller's max_output_bytes                  */
    EDS_ERR_HALO = 7,             /* a symbol cannot be resolved inside the shard's window: widen it   */
    EDS_ERR_BAD_VCF = 8           /* VCF/FASTA input outside the reference's well-defined domain (C.4) */
} eds_status;

typedef struct eds_ctx eds_ctx; /* one per (device, stream); owns scratch + output buffers; not thread-safe */

/* Thread-local message of the last failing call on this thread. */
const char* eds_last_error(void);
/* "edsparser_b200 <version> sm_100a" */
const char* eds_version(void);

/* stream: a cudaStream_t to launch on (e.g. the caller's current stream), or NULL = the library
 * creates its own non-blocking stream. */
eds_status eds_ctx_create(int device, void* stream, eds_ctx** out);
void eds_ctx_destroy(eds_ctx* ctx);
eds_status eds_ctx_synchronize(eds_ctx* ctx);

/* Launch-shape knobs (tests sweep them; results never depend on them).
 * partitions: blocks used by the scan/compaction kernels (0 = default). */
eds_status eds_ctx_set_tuning(eds_ctx* ctx, uint32_t partitions, uint32_t scan_blocks_per_sm);

/* Per-kernel device times of the LAST transform on this ctx, measured with CUDA events on the
 * ctx's stream when profiling is on (adds one event pair per launch; leave off for throughput runs).
 * eds_ctx_kernel_times fills up to cap entries and returns the number of kernels launched. */
eds_status eds_ctx_set_profiling(eds_ctx* ctx, int on);
uint32_t eds_ctx_kernel_times(eds_ctx* ctx, const char** names, float* ms, uint32_t cap);

/* ------------------------------------------------------------------------------------------
 * MSA loader (host side): locate the rows of a gapped-FASTA alignment.
 * Replaces the bookkeeping half of parse_msa_and_build_variant_bv (msa_transforms.cpp:36-90):
 * MSAMetadata{start_positions, n_sequences, seq_length, line_width} (msa_transforms.cpp:18-24).
 * Cost is O(first row + rows x header length); the byte comparison itself runs on the GPU.
 * ------------------------------------------------------------------------------------------ */
typedef struct eds_msa_index {
    uint64_t* row_start; /* malloc'd, n_rows entries: offset of the first residue of each row */
    uint32_t n_rows;     /* R */
    uint64_t n_cols;     /* C (residues per row, gaps included) */
    uint32_t line_width; /* residues per text line */
    uint64_t row_bytes;  /* bytes from a row's first to last residue, inner newlines included */
} eds_msa_index;

eds_status eds_msa_index_host(const uint8_t* text, uint64_t text_bytes, eds_msa_index* out);
void eds_msa_index_free(eds_msa_index* idx);

/* A column window of an alignment held in ONE device buffer. Residue (r, c),
 * col_begin <= c < col_begin + col_count, lives at
 *     text[row_start[r] + (c + c / line_width) - (col_begin + col_begin / line_width)]
 * i.e. each row segment is a verbatim slice of the FASTA text, inner '\n' included, rows may start
 * at any byte offset. `text` itself must be 16-byte aligned and readable up to the next 16-byte
 * boundary past text_bytes (any cudaMalloc'd buffer is).
 * A whole file is the window col_begin = 0, col_count = total_cols, own = [0, total_cols).
 * A shard owns the symbols that START in [own_begin, own_end) plus the conserved text of its own
 * columns, and holds a halo on both sides: >= 1 column (EDS) or >= l + 1 columns (l-EDS) on a side
 * that is not an end of the alignment, and far enough to the right to contain the end of its last
 * variable symbol (else EDS_ERR_HALO). Concatenating the shards' outputs in column order gives the
 * whole-alignment output byte for byte. col_count < 2^31 - 64. */
typedef struct eds_msa_view {
    const uint8_t* text;
    uint64_t text_bytes;
    const uint64_t* row_start; /* HOST array, n_rows entries */
    uint32_t n_rows;
    uint32_t line_width;
    uint64_t total_cols;
    uint64_t col_begin, col_count;
    uint64_t own_begin, own_end;
} eds_msa_view;

typedef struct eds_buffer {
    uint8_t* data; /* device memory owned by the ctx (valid until the next transform on it), or malloc'd host memory */
    uint64_t bytes;
} eds_buffer;

typedef struct eds_msa_stats {
    uint64_t n_variable_cols; /* non-conserved columns in the window                              */
    uint64_t n_runs;          /* maximal runs of conserved / variable columns in the window       */
    uint64_t n_symbols;       /* symbols that start in the owned range                            */
    uint64_t n_variable;      /* ... of which variable (degenerate or single-haplotype)           */
    uint64_t n_alternatives;  /* strings in those variable symbols                                */
    uint64_t first_open_col;  /* global column of the first owned symbol (UINT64_MAX if none)     */
    uint64_t eds_bytes, seds_bytes; /* bytes this shard emits                                     */
    uint64_t eds_lead_bytes;  /* of which continue a conserved symbol opened by a lower shard      */
    uint32_t tail_open_common; /* 1: the last owned symbol is conserved and continues in the next shard */
    uint32_t gpu_launches;    /* kernels launched by this call (retries included)                 */
    uint32_t retries;         /* pipeline re-runs after a scratch/output buffer had to grow       */
    uint32_t n_hashed_symbols; /* multi-column symbols that took the hashed row path (outside the tuple form's envelope) */
} eds_msa_stats;

/* Whole pipeline on one device, input already in device memory.
 * leds = 0: parse_msa_to_eds_streaming  (msa_transforms.cpp:334-345)
 * leds = 1: parse_msa_to_leds_streaming (msa_transforms.cpp:351-365) with context_length = l
 * Outputs are byte-identical to the reference's pair<string,string> (for a shard: to this shard's
 * slice of it); they stay in device memory owned by the ctx. The call returns after the stream
 * has been synchronised (the output sizes are needed on the host). */
eds_status eds_msa_transform_device(eds_ctx* ctx, const eds_msa_view* view, uint32_t l, int leds,
                                    eds_buffer* eds_out, eds_buffer* seds_out, eds_msa_stats* stats);

/* Same, from the bytes of a .msa file in host memory (pinned memory copies fastest) to malloc'd
 * host strings: index on the host, H2D, transform, D2H. This is what the C++ wrappers and the CLI
 * call. Free the outputs with eds_buffer_free_host. */
eds_status eds_msa_transform_host(eds_ctx* ctx, const uint8_t* file, uint64_t file_bytes, uint32_t l, int leds,
                                  eds_buffer* eds_out, eds_buffer* seds_out, eds_msa_stats* stats);

/* Same call; the results are VIEWS into pinned host memory owned by the ctx (valid until the next *_view call on this
 * ctx, never to be freed by the caller): no allocation and no first-touch page faults per call. */
eds_status eds_msa_transform_host_view(eds_ctx* ctx, const uint8_t* file, uint64_t file_bytes, uint32_t l, int leds,
                                       eds_buffer* eds_out, eds_buffer* seds_out, eds_msa_stats* stats);

/* Conserved-column bit vector B of msa_transforms.cpp:36-90 for the window: out_bits[c / 8] bit
 * (c % 8) = 1 iff window column c is conserved; out_bytes >= ceil(col_count / 8). For tests. */
eds_status eds_msa_conserved_bits(eds_ctx* ctx, const eds_msa_view* view, uint8_t* out_bits, uint64_t out_bytes);

/* Synthetic alignment of BASELINE.json configs 2 and 4, generated directly in
 * device memory as FASTA text for the column window [col_begin, col_begin + col_count):
 * headers ">seq<r+1>", wrap `line_width`, counter-based RNG keyed on (seed, row, column) so any
 * window of the same (seed, R, C) alignment agrees with the whole. The buffer is owned by the ctx
 * until eds_msa_synth_free / ctx destruction; `view` is filled in (own range = the whole window;
 * callers narrow it), view->row_start points at ctx-owned host memory.
 * variable_ppm: probability (per million) that a column is variable. */
eds_status eds_msa_synth_device(eds_ctx* ctx, uint32_t n_rows, uint64_t total_cols, uint32_t line_width,
                                uint64_t col_begin, uint64_t col_count, uint64_t seed, uint32_t variable_ppm,
                                eds_msa_view* view);
void eds_msa_synth_free(eds_ctx* ctx);

/* Copy a device-resident output (eds_msa_transform_device) into malloc'd host memory. */
eds_status eds_buffer_to_host(eds_ctx* ctx, const eds_buffer* device_buf, eds_buffer* host_out);
void eds_buffer_free_host(eds_buffer* buf);
/* Same copy into pinned host memory kept by the ctx (slot 0 or 1, grow-only): host_out is a VIEW, valid until the
 * next *_view call that uses the slot, never to be freed by the caller. */
eds_status eds_buffer_to_host_view(eds_ctx* ctx, int slot, const eds_buffer* device_buf, eds_buffer* host_out);

/* ------------------------------------------------------------------------------------------
 * l-EDS merge: eds_to_leds_linear (eds_transforms.cpp:313-373) when seds_in != NULL,
 * eds_to_leds_cartesian (eds_transforms.cpp:381-426) when seds_in == NULL.
 * Host text in, malloc'd host text out (EDS::save / save_sources dialect, eds.cpp:600-659).
 * max_output_bytes = 0 means no limit (the reference has none); otherwise EDS_ERR_BUDGET.
 * ------------------------------------------------------------------------------------------ */
eds_status eds_leds_merge_host(eds_ctx* ctx, const uint8_t* eds_in, uint64_t eds_bytes, const uint8_t* seds_in,
                               uint64_t seds_bytes, uint32_t l, int compact, uint64_t max_output_bytes,
                               eds_buffer* leds_out, eds_buffer* seds_out, uint32_t* rounds_out);

/* Same merge; the results are VIEWS into pinned host memory owned by the ctx (shared with
 * eds_vcf_transform_host_view): valid until the next *_view call on this ctx, never to be freed by the caller. */
eds_status eds_leds_merge_host_view(eds_ctx* ctx, const uint8_t* eds_in, uint64_t eds_bytes, const uint8_t* seds_in,
                                    uint64_t seds_bytes, uint32_t l, int compact, uint64_t max_output_bytes,
                                    eds_buffer* leds_out, eds_buffer* seds_out, uint32_t* rounds_out);

/* The same merge with the inputs ALREADY in device memory (16-byte aligned, e.g. eds_genrandomeds_device's output):
 * host text out. */
eds_status eds_leds_merge_device_in(eds_ctx* ctx, const uint8_t* eds_dev, uint64_t eds_bytes, const uint8_t* seds_dev,
                                    uint64_t seds_bytes, uint32_t l, int compact, uint64_t max_output_bytes,
                                    eds_buffer* leds_out, eds_buffer* seds_out, uint32_t* rounds_out);

/* genrandomeds on the device (reference tool src/cpp/tools/genrandomeds.cpp:221-352; BASELINE config 3's generator):
 * an EDS + SEDS pair of the tool's shape — n random bases, a position is a variant site with probability
 * variability_ppm / 10^6, 2..4 alternatives per site (reference base first; 70 % SNP, 15 % insertion of 1..10 bases,
 * 15 % empty), `paths` paths, conserved runs carry {0}, no trailing newline — written straight into device memory
 * owned by the ctx (valid until the next call). Keyed on (seed, position) with a counter-based generator, not the
 * tool's sequential mt19937 stream: shape-compatible, not stream-compatible. */
eds_status eds_genrandomeds_device(eds_ctx* ctx, uint64_t ref_size, uint32_t variability_ppm, uint32_t paths, uint64_t seed,
                                   eds_buffer* eds_out, eds_buffer* seds_out);

/* is_leds (eds_transforms.cpp:439-468): 1 iff no interior non-degenerate symbol is shorter than l and no
 * two degenerate symbols are adjacent (l = 0: always 1). Parsed and tested on the device. */
eds_status eds_is_leds_host(eds_ctx* ctx, const uint8_t* eds_in, uint64_t eds_bytes, uint32_t l, int* is_leds_out);

/* ------------------------------------------------------------------------------------------
 * EDS objects: EDS::parse + normalize_eds_format (eds.cpp:39-155, 831-881), EDS::parse_sources
 * (eds.cpp:268-355), calculate_statistics / calculate_source_statistics (eds.cpp:361-505) and
 * EDS::merge_adjacent (eds.cpp:1425-1695), run on the device. The host class `edsparser::EDS`
 * (edsparser_b200/host/include/edsparser/formats/eds.hpp) is built from what these return.
 * ------------------------------------------------------------------------------------------ */
typedef struct eds_parsed {
    uint8_t* text;        /* the EDS text without white space (compact input stays compact), malloc'd       */
    uint64_t text_bytes;
    uint32_t* str_start;  /* n_strings: string j is text[str_start[j] .. str_end[j])                        */
    uint32_t* str_end;
    uint32_t* sym_first;  /* n_symbols + 1: symbol i holds strings sym_first[i] .. sym_first[i + 1)          */
    uint32_t n_strings, n_symbols;
    uint32_t has_sources;
    uint64_t* src_off;    /* n_strings + 1 (sources given): ids of string j are src_ids[src_off[j] ..)      */
    int32_t* src_ids;     /* ascending, duplicates removed (std::set<int> order)                            */
    /* EDS::Statistics (eds.hpp:105-119), reduced on the device */
    uint64_t total_chars;            /* N */
    uint64_t num_degenerate_symbols, num_common_chars, total_change_size, num_empty_strings;
    uint64_t sum_context_length, num_context_blocks;
    uint32_t min_context_length, max_context_length;
    uint64_t num_paths, max_paths_per_string, total_paths;
} eds_parsed;

/* seds == NULL: no sources. Errors carry the reference's messages (runtime_error / out_of_range). */
eds_status eds_parse_host(eds_ctx* ctx, const uint8_t* eds, uint64_t eds_bytes, const uint8_t* seds, uint64_t seds_bytes,
                          eds_parsed* out);
void eds_parsed_free(eds_parsed* parsed);

/* EDS::merge_adjacent(pos1, pos1 + 1): all combinations without sources, the combinations with a non-empty source
 * intersection (a set containing 0 is universal) with them. Output in the FULL dialect of EDS::save /
 * save_sources (trailing newline), malloc'd. The caller checks adjacency / range (the reference's messages name
 * both positions). */
eds_status eds_merge_adjacent_host(eds_ctx* ctx, const uint8_t* eds, uint64_t eds_bytes, const uint8_t* seds,
                                   uint64_t seds_bytes, uint64_t pos1, eds_buffer* eds_out, eds_buffer* seds_out);

/* ------------------------------------------------------------------------------------------
 * VCF front end: parse_vcf_to_eds_streaming (vcf_transforms.cpp:677-729) when l == 0,
 * parse_vcf_to_leds_streaming (vcf_transforms.cpp:735-755: the same followed by the LINEAR merge,
 * one thread, compact) when l > 0. One path id per sample column, 1-based (vcf_transforms.cpp:588).
 * Only the first FASTA record is used and CHROM is ignored, as in the reference.
 * Inputs the reference is undefined on return EDS_ERR_BAD_VCF instead of garbage: POS 0, a record
 * that extends past the end of the FASTA sequence, FASTA lines of unequal width, '\r' in the FASTA.
 * ------------------------------------------------------------------------------------------ */
typedef struct eds_vcf_stats {
    /* VCFStats (vcf_transforms.hpp:24-36) */
    uint64_t total_variants, processed_variants, skipped_malformed, skipped_unsupported_sv, variant_groups;
    /* this implementation */
    uint64_t n_lines;           /* lines of the VCF text, headers included                          */
    uint64_t n_bases;           /* length of the first FASTA record                                 */
    uint64_t n_alleles;         /* carrier bitsets built: one per (record, allele incl. REF)        */
    uint64_t n_haplotype_slots; /* haplotype candidates over all groups before de-duplication       */
    uint64_t n_samples_max;     /* widest sample matrix row                                         */
    uint64_t eds_bytes, seds_bytes; /* size of the EDS / SEDS text before the optional l-EDS merge  */
    uint32_t gpu_launches;      /* kernels launched by the front end                                */
    uint32_t host_sorted;       /* 1: positions were not strictly increasing; the tie order of the
                                   reference's unstable std::sort was reproduced on the host        */
    uint32_t retries;           /* re-runs of the genotype kernel after the bitset width grew       */
    uint32_t leds_rounds;       /* merge rounds (l > 0)                                             */
} eds_vcf_stats;

/* Host text in, malloc'd host text out (free with eds_buffer_free_host). sv_lines (optional):
 * malloc'd byte offsets, in file order, of the records skipped for an unsupported symbolic ALT
 * (vcf_transforms.cpp:299-305; the caller prints the reference's warning from them); free(). */
eds_status eds_vcf_transform_host(eds_ctx* ctx, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta,
                                  uint64_t fasta_bytes, uint32_t l, eds_buffer* eds_out, eds_buffer* seds_out,
                                  eds_vcf_stats* stats, uint64_t** sv_lines, uint64_t* n_sv_lines);

/* Same call, but the results are VIEWS into pinned host memory owned by the ctx (grow-only, reused by every call):
 * valid until the next eds_vcf_transform_host_view on this ctx, never to be freed by the caller. A fresh malloc'd
 * gigabyte costs more in first-touch page faults and pageable copies than the whole device pipeline; a caller that
 * streams the text on (writes the files, feeds a socket) should use this form. */
eds_status eds_vcf_transform_host_view(eds_ctx* ctx, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta,
                                       uint64_t fasta_bytes, uint32_t l, eds_buffer* eds_out, eds_buffer* seds_out,
                                       eds_vcf_stats* stats, uint64_t** sv_lines, uint64_t* n_sv_lines);

/* l == 0 only, everything in device memory: both inputs 16-byte aligned and readable up to the next
 * 16-byte boundary (any cudaMalloc'd buffer is); outputs are owned by the ctx until its next VCF call. */
eds_status eds_vcf_transform_device(eds_ctx* ctx, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta,
                                    uint64_t fasta_bytes, eds_buffer* eds_out, eds_buffer* seds_out,
                                    eds_vcf_stats* stats);


/* ------------------------------------------------------------------------------------------
 * Multi-GPU msa2eds: the alignment's columns shard across the GPUs of one node (SURVEY.md §8e,
 * BASELINE config 4). Every device transforms a contiguous column range plus a halo (widened and
 * retried on EDS_ERR_HALO); the ONE exchange is an NCCL all-gather of the shards' (eds, seds) byte
 * counts, from which every shard gets the offsets of its slices of the single output pair. NCCL is
 * called from this library (dlopen of libnccl.so.2 on first use), not from the caller.
 *
 * eds_group: ONE process drives N devices (what `msa2eds --gpus N` uses): one context and one host
 * thread per device, an in-process communicator. Stands behind parse_msa_to_eds_streaming /
 * parse_msa_to_leds_streaming (msa_transforms.cpp:334-365) like the single-device entry points.
 * ------------------------------------------------------------------------------------------ */
typedef struct eds_group eds_group;
/* devices = NULL: devices 0 .. n_devices - 1 */
eds_status eds_group_create(const int* devices, int n_devices, eds_group** out);
void eds_group_destroy(eds_group* group);
int eds_group_size(const eds_group* group);
eds_ctx* eds_group_ctx(eds_group* group, int i); /* e.g. for eds_ctx_set_profiling / eds_ctx_kernel_times */
/* .msa bytes in host memory -> malloc'd host strings (free with eds_buffer_free_host). halo = 0: default (4096
 * columns, never less than l + 2). per_device_stats: NULL or eds_group_size() entries. */
eds_status eds_group_msa_transform_host(eds_group* group, const uint8_t* file, uint64_t file_bytes, uint32_t l, int leds,
                                        uint64_t halo, eds_buffer* eds_out, eds_buffer* seds_out,
                                        eds_msa_stats* per_device_stats);
/* Same, every device pwrite()s its slices into the two open files at its offsets (device 0 sizes them);
 * totals[0..1] = bytes of the .eds / .seds. */
eds_status eds_group_msa_transform_fd(eds_group* group, const uint8_t* file, uint64_t file_bytes, uint32_t l, int leds,
                                      uint64_t halo, int eds_fd, int seds_fd, uint64_t totals[2],
                                      eds_msa_stats* per_device_stats);

/* eds2leds over the devices of a group (symbol ranges): the EDS text is cut INSIDE long conserved
 * symbols (no candidate pair of select_independent_merge_pairs, eds_transforms.cpp:46-107, can involve a conserved
 * symbol of >= l characters unless a neighbour collapses to a short conserved one), every shard is merged on its own
 * device, the halves are glued back together. Every shard reports whether its seam symbols stayed unmerged; if one did
 * not, or the text offers no such symbol near a cut, the whole text is merged on one device. The bytes are those of
 * eds_leds_merge_host either way; *shards_used tells which way it went. seds = NULL: CARTESIAN. */
eds_status eds_group_leds_merge_host(eds_group* group, const uint8_t* eds, uint64_t eds_bytes, const uint8_t* seds,
                                     uint64_t seds_bytes, uint32_t l, int compact, eds_buffer* leds_out,
                                     eds_buffer* seds_out, uint32_t* rounds_out, uint32_t* shards_used);

/* vcf2eds over the devices of a group: slices of the VCF's record lines, one per device, each
 * with the whole FASTA record; the slices meet once inside the transform (one sort of all positions for the tie order
 * of the reference's unstable std::sort, vcf_transforms.cpp:715-718; no overlapping group may span a cut; each slice
 * renders the reference bases up to the next slice's first group). Arguments and results as eds_vcf_transform_host
 * (vcf_transforms.cpp:677-755); l > 0: the slices are joined in the first device's memory (peer copies) and merged there. Input the slices cannot be
 * joined on (a slice without records, a spanning group, an error) runs on the group's first device.
 * shards_used (optional): devices that took part. */
eds_status eds_group_vcf_transform_host(eds_group* group, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta,
                                        uint64_t fasta_bytes, uint32_t l, eds_buffer* eds_out, eds_buffer* seds_out,
                                        eds_vcf_stats* stats, uint64_t** sv_lines, uint64_t* n_sv_lines,
                                        uint32_t* shards_used);

/* Same call; the results are VIEWS into pinned host memory (owned by the group when the slices were joined, by the first
 * device's context when the input ran whole there): valid until the next *_view call on this group, never freed by the
 * caller. Every device copies its slice over its own PCIe link straight into the one pinned pair. */
eds_status eds_group_vcf_transform_host_view(eds_group* group, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta,
                                             uint64_t fasta_bytes, uint32_t l, eds_buffer* eds_out, eds_buffer* seds_out,
                                             eds_vcf_stats* stats, uint64_t** sv_lines, uint64_t* n_sv_lines,
                                             uint32_t* shards_used);

/* eds_comm: one process PER GPU (torchrun, mpirun): rank 0 makes the 128-byte NCCL id, the launcher ships it to
 * every rank, each rank builds the communicator for its context. After every eds_msa_transform_device the rank
 * posts its byte counts (host values by then: the exchange runs on the comm's own stream beside the next transform,
 * no host synchronisation, four posts may be in flight); eds_comm_offsets waits for the last post and returns
 * {eds offset, seds offset, eds total, seds total}; eds_comm_flush makes the context's stream wait for it. */
typedef struct eds_comm eds_comm;
eds_status eds_nccl_unique_id(uint8_t out[128]);
eds_status eds_comm_create(eds_ctx* ctx, const uint8_t id[128], int rank, int world, eds_comm** out);
void eds_comm_destroy(eds_comm* comm);
eds_status eds_comm_post(eds_comm* comm, uint64_t eds_bytes, uint64_t seds_bytes);
eds_status eds_comm_offsets(eds_comm* comm, uint64_t out[4]);
eds_status eds_comm_flush(eds_comm* comm);

/* Device memory for callers that stage inputs themselves (bench, tests): cudaMalloc / H2D copy / cudaFree. */
eds_status eds_device_upload(eds_ctx* ctx, const uint8_t* host, uint64_t bytes, uint8_t** device_out);
void eds_device_free(eds_ctx* ctx, uint8_t* device);

#ifdef __cplusplus
}
#endif
#endif /* EDSPARSER_B200_H */
