/*
 * radian_b200 -- C ABI of the B200-native RADIAN decode hot path.
 *
 * The reference (comprna/radian) has no FFI layer: its hot path is two in-process Python
 * functions.  These entry points are what a binding for that path calls instead:
 *
 *   radian_decode_batch_*    replaces  decode.beam_search            radian/decode.py:100-212
 *                            (call sites radian/basecall.py:102-109 global, :113-120 chunk)
 *   radian_assemble_batch_*  replaces  matrix_assembly.assemble_matrices
 *                                                                    radian/matrix_assembly.py:6-53
 *                            (call site radian/basecall.py:100)
 *   radian_stitch_batch_*    replaces  sequence_assembly.simple_assembly + argmax + index2base
 *                                                                    radian/sequence_assembly.py:19-48,90-97
 *                            (call site radian/basecall.py:122-123)
 *   radian_normalise_batch_* replaces  preprocess.mad_normalise         radian/preprocess.py:23-49
 *   radian_windows_*         replaces  preprocess.get_windows           radian/preprocess.py:4-21
 *                            (call sites radian/basecall.py:78,83)
 *   radian_table_*           replaces  the dict built from the RNA model JSON
 *                                                                    radian/basecall.py:47-57
 *                            and the entropy memo `entr_cache`       radian/decode.py:86-90
 *
 * Conventions
 *   - every function returns 0 on success, a negative RADIAN_E_* code otherwise;
 *     radian_last_error() gives a thread-local message for the last failure.
 *   - symbols: A,C,G,T = 0..3, CTC blank = column 4 of every posterior row (decode.py:124).
 *   - sequences are returned in decode order (3'->5'); the caller reverses them, exactly as
 *     basecall.py:129 does.
 *   - "_dev" functions take device pointers and enqueue work on `stream` without
 *     synchronising; "_host" functions take host pointers, copy in, run, copy out and
 *     synchronise before returning.
 *   - the library never falls back to a CPU implementation: without a CUDA device every
 *     compute entry point fails with RADIAN_E_CUDA.
 */
#ifndef RADIAN_B200_H
#define RADIAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RADIAN_OK 0
#define RADIAN_E_ARG (-1)      /* bad argument (ValueError in the Python layer) */
#define RADIAN_E_CUDA (-2)     /* CUDA runtime failure or no device */
#define RADIAN_E_CONTEXT (-3)  /* len_context does not match the table (KeyError, decode.py:83) */
#define RADIAN_E_READ (-4)     /* at least one read failed; see out_status */
#define RADIAN_E_GAP (-5)      /* a chunk starts past the rows assembled so far (IndexError,
                                  matrix_assembly.py:27) */

/* per-read status written to out_status */
#define RADIAN_READ_OK 0
#define RADIAN_READ_SEQ_OVERFLOW 1   /* decoded sequence longer than the caller's slot */
#define RADIAN_READ_TRIE_OVERFLOW 2  /* back-pointer arena too small even after compaction */
#define RADIAN_READ_INDEX_ERROR 3    /* stitching: a fragment does not fit the reference's vote buffer
                                       (IndexError in add_count, sequence_assembly.py:47) */

#define RADIAN_READ_EMPTY_SIGNAL 4   /* preprocessing: ValueError("Signal must not be empty to normalise"),
                                       preprocess.py:24-25 */
#define RADIAN_READ_MAD_ZERO 5       /* preprocessing: ValueError("MAD is zero, issue with signal."),
                                       preprocess.py:47-48 */

#define RADIAN_READ_RANGE 6          /* decode: two kept beams further apart than the float64 exponent range
                                       lets the kernel's rescaled linear-domain scores express (the
                                       reference's log-domain scores, decode.py:172-175, have no such
                                       limit); never happens for posteriors above ~1e-150 */

#define RADIAN_READ_KEY_ERROR 7      /* decode: a kept beam reached a context that the model does not hold
                                       (KeyError at lm[context], decode.py:83); out_len holds the index of
                                       that context.  Only tables made by radian_table_create_sparse. */

#define RADIAN_MAX_BEAM_WIDTH 128
#define RADIAN_MAX_CONTEXT 13

typedef struct radian_table radian_table_t;
typedef void *radian_stream_t; /* cudaStream_t */

const char *radian_last_error(void);
const char *radian_version(void);
int radian_device_count(void);
/* The _host entry points keep their device buffers between calls in a stream-ordered memory pool
 * that belongs to this library (never the device's default pool); this returns them to the driver. */
int radian_trim_memory(int device);

/*
 * Dense RNA k-mer table.  probs: host array of 4^L rows x 4 float64 linear probabilities,
 * row index = big-endian base-4 value of the L context symbols, oldest symbol most
 * significant (the key order of basecall.py:54-57).  The row entropies of decode.py:86-90
 * are computed once here, on the host, in float64 with libm's log in the reference's
 * operation order, and uploaded with the rows; the table stays resident in HBM.
 */
int radian_table_create(const double *probs, int L, int device, radian_table_t **out);
/* The same for a model that does not hold every context (the reference's dict may be sparse: it fails
 * with KeyError only when the search reaches a missing context, decode.py:83).  present: 4^L bytes,
 * 0 = context absent (its row in probs is ignored); NULL = all present. */
int radian_table_create_sparse(const double *probs, const uint8_t *present, int L, int device,
                               radian_table_t **out);
int radian_table_destroy(radian_table_t *t);
int radian_table_context_len(const radian_table_t *t);
/* copies the 4^L float64 row entropies back to the host (tests) */
int radian_table_entropies(const radian_table_t *t, double *out_host);

/*
 * CTC prefix beam search with optional gated RNA-model fusion over a batch of reads.
 *
 *  post          concatenated posterior rows, 5 per frame, float32 (post_is_f64 == 0) or
 *                float64; read r owns rows [frame_offsets[r], frame_offsets[r+1]).
 *  order         optional permutation of 0..n_reads-1: the order in which reads are handed to
 *                the device work queue (longest first balances the tail); NULL = 0,1,2,...
 *  beam_width    1..RADIAN_MAX_BEAM_WIDTH (decode.py:145).
 *  table         NULL = RNA model off (lm falsy, decode.py:157,180).  When given, len_context
 *                must equal radian_table_context_len(table).
 *  s_threshold, r_threshold   decode.py:93, both comparisons strict.
 *  out_seq       symbols of the best labeling of read r at out_seq[seq_offsets[r] ...];
 *                slot size seq_offsets[r+1]-seq_offsets[r] (T_r is always enough).
 *  out_len       decoded length per read.
 *  out_score     2 per read: natural-log pr_total of the best and of the second-best final
 *                beam (-inf when it has probability 0, NaN when there is no second beam).
 *  out_status    RADIAN_READ_* per read.
 *  out_counters  optional, 4 per read: number of lm[context] reads the reference would have
 *                done (decode.py:83), number of combine_dists calls (decode.py:94), number of
 *                frames whose selection ranked two candidates within 2^-40 of each other (the
 *                reference decides those on the rounding noise of its log-domain sums; a read
 *                whose sequence differs from the reference's has a non-zero count here and, for
 *                a tie of the final beams, equal out_score entries), and a diagnostic word: high 32
 *                bits = frames whose signal entropy gate was open (H_s > s_threshold, decode.py:93),
 *                low 32 bits = frames in which the kernel had to look at extensions at all (the
 *                others only updated the scores of the kept beams).
 */
/*
 *  max_frames    frames of the longest read; total_frames: frames of all reads (= frame_offsets[n_reads],
 *                which lives in device memory), or 0 if the caller does not know it.  Both only shape
 *                the launch: a batch that does not fill the GPU several times over runs with fewer
 *                resident warps per SM, which finishes its longest read sooner.
 *  arena_nodes   capacity of the per-read back-pointer arena; 0 picks a default from max_frames
 *                (exact worst case for small problems, else beam lanes x max_frames/32).  A read
 *                that needs more gets RADIAN_READ_TRIE_OVERFLOW; lanes x (T+1) always suffices.
 *                The _host entry point retries such reads by itself.
 */
size_t radian_decode_workspace_bytes(int device, int beam_width, int n_reads, int64_t max_frames,
                                     int64_t arena_nodes);

int radian_decode_batch_dev(const void *post, int post_is_f64, const int64_t *frame_offsets,
                            int n_reads, const int32_t *order, int64_t max_frames,
                            int64_t total_frames, int beam_width,
                            const radian_table_t *table, int len_context, double s_threshold,
                            double r_threshold, uint8_t *out_seq, const int64_t *seq_offsets,
                            int64_t *out_len, double *out_score, int32_t *out_status,
                            uint64_t *out_counters, int64_t arena_nodes, void *workspace,
                            size_t workspace_bytes, radian_stream_t stream);

int radian_decode_batch_host(const void *post, int post_is_f64, const int64_t *frame_offsets,
                             int n_reads, int beam_width, const radian_table_t *table,
                             int len_context, double s_threshold, double r_threshold,
                             uint8_t *out_seq, const int64_t *seq_offsets, int64_t *out_len,
                             double *out_score, int32_t *out_status, uint64_t *out_counters,
                             int device);

/* The same for reads that are separate matrices in host memory (what the reference's callers hold:
 * one numpy array per read or per window, basecall.py:99-120): reads[r] points at read r's
 * n_frames[r] x 5 values.  Nothing is concatenated on the host; the reads are gathered straight into
 * the transfer. */
int radian_decode_batch_host_reads(const void *const *reads, const int64_t *n_frames, int post_is_f64,
                                   int n_reads, int beam_width, const radian_table_t *table,
                                   int len_context, double s_threshold, double r_threshold,
                                   uint8_t *out_seq, const int64_t *seq_offsets, int64_t *out_len,
                                   double *out_score, int32_t *out_status, uint64_t *out_counters,
                                   int device);

/*
 * Merge of overlapping chunk posteriors for a batch of reads (matrix_assembly.py:6-53,
 * including its first-chunk-wins behaviour: np.add's result is discarded at :52).
 *
 *  chunks             all chunk rows of all reads back to back, 5 float32 per row.
 *  chunk_row_offsets  n_chunks+1 row offsets into `chunks`.
 *  read_chunk_ranges  n_reads+1 chunk indices: read r owns chunks
 *                     [read_chunk_ranges[r], read_chunk_ranges[r+1]), the k-th of them starts at
 *                     global row k*step of the read (create_vstack, :12-34).
 *  out_row_offsets    n_reads+1 row offsets into `out` (row counts from radian_assemble_plan).
 *  max_chunk_rows     as returned by radian_assemble_plan: the largest chunk, negated when the
 *                     layout is ragged (some chunk other than a read's last is shorter), which
 *                     selects the general lookup instead of the closed form.
 *  out_is_f64         per batch: 1 writes float64 rows (rows covered by more than one chunk are
 *                     L1-normalised in float64, others are exact casts), 0 writes float32
 *                     (only valid when no row of the batch is covered twice).
 */
int radian_assemble_plan(const int64_t *chunk_row_offsets, const int64_t *read_chunk_ranges,
                         int n_reads, int step, int64_t *out_rows_per_read, int *out_any_overlap,
                         int32_t *out_max_chunk_rows);

int radian_assemble_batch_dev(const float *chunks, const int64_t *chunk_row_offsets,
                              const int64_t *read_chunk_ranges, const int64_t *out_row_offsets,
                              int n_reads, int step, int32_t max_chunk_rows,
                              int64_t total_out_rows, void *out, int out_is_f64,
                              radian_stream_t stream);

int radian_assemble_batch_host(const float *chunks, const int64_t *chunk_row_offsets,
                               const int64_t *read_chunk_ranges, const int64_t *out_row_offsets,
                               int n_reads, int step, void *out, int out_is_f64, int device);

/*
 * Chunk-mode stitching for a batch of reads: replaces sequence_assembly.simple_assembly
 * (radian/sequence_assembly.py:19-48) followed by np.argmax(consensus, axis=0) and index2base
 * (:90-97), call site radian/basecall.py:122-123.  Consecutive fragments of a read are aligned
 * on the first largest matching block of difflib.SequenceMatcher(None, previous, current)
 * (including its autojunk rule for fragments of 200 symbols and more), every fragment votes for
 * its symbols at its position, and every consensus column takes the first maximum.
 *
 *  frag_sym          symbols 0..3 (A,C,G,T) of all fragments of all reads back to back, i.e. the
 *                    out_seq of a chunk-mode radian_decode_batch_* call.
 *  frag_offsets      n_frags+1 offsets into frag_sym.
 *  read_frag_ranges  n_reads+1 fragment indices: read r owns fragments
 *                    [read_frag_ranges[r], read_frag_ranges[r+1]) in chunk order.
 *  out_seq           consensus symbols of read r at out_seq[out_offsets[r] ...]; a slot as large as
 *                    the sum of the read's fragment lengths always suffices.
 *  out_len           consensus length per read.  As in the reference it only counts columns
 *                    reached by the second and later fragments: a read with a single fragment
 *                    yields an empty consensus (sequence_assembly.py:39).
 *  out_status        RADIAN_READ_OK, or RADIAN_READ_INDEX_ERROR where the reference raises
 *                    IndexError (the call then returns RADIAN_E_GAP).
 *  out_votes         optional: 4 int32 vote counts (A,C,G,T) per consensus column, at
 *                    out_votes[4 * (out_offsets[r] + column)].
 */
int radian_stitch_batch_host(const uint8_t *frag_sym, const int64_t *frag_offsets,
                             const int64_t *read_frag_ranges, int n_reads, uint8_t *out_seq,
                             const int64_t *out_offsets, int64_t *out_len, int32_t *out_status,
                             int32_t *out_votes, int device);

/*
 * The same on device pointers, asynchronous on `stream`.  Fragments are given by start and length
 * into frag_sym, so the out_seq / seq_offsets / out_len of a chunk-mode radian_decode_batch_dev call
 * can be passed as they are (frag_start = seq_offsets, frag_len = out_len).  out_offsets are the
 * caller's slots (total_slots = out_offsets[n_reads]; a read whose consensus does not fit its slot
 * gets RADIAN_READ_SEQ_OVERFLOW).  Pairs whose second fragment has 200 symbols or more run
 * difflib's full recursion in a scratch pool of long_pair_scratch_ints 32-bit words inside the
 * workspace (0 if no fragment is that long; a read that finds the pool empty gets
 * RADIAN_READ_TRIE_OVERFLOW).  Statuses are per read; the call itself returns RADIAN_OK.
 */
size_t radian_stitch_workspace_bytes(int64_t n_frags, int64_t total_slots, int64_t long_pair_scratch_ints);
int radian_stitch_batch_dev(const uint8_t *frag_sym, const int64_t *frag_start, const int64_t *frag_len,
                            const int64_t *read_frag_ranges, int n_reads, int64_t n_frags, uint8_t *out_seq,
                            const int64_t *out_offsets, int64_t total_slots, int64_t *out_len,
                            int32_t *out_status, int32_t *out_votes, int64_t long_pair_scratch_ints,
                            void *workspace, size_t workspace_bytes, radian_stream_t stream);

/*
 * Signal preprocessing for a batch of reads: replaces preprocess.mad_normalise
 * (radian/preprocess.py:23-49, call site radian/basecall.py:78).
 *
 *  signal           raw int16 samples of all reads back to back (fast5 Raw/Signal); read r owns
 *                   [offsets[r], offsets[r+1]).
 *  outlier_z_score  --outlier-clip; outlier_is_int says whether the caller's value is an integer
 *                   (argparse type=int, basecall.py:25), which matters for the result type below.
 *  out              8 bytes per sample at the same offsets: float64 modified z-scores
 *                   (x - median) / (1.4826 * MAD) clipped to +-outlier_z_score, or, where
 *                   out_is_int64[r] is set, int64 values truncated towards zero: np.vectorize takes
 *                   its output type from the first element, and a clipped first sample returns the
 *                   integer outlier_z_score itself.
 *  out_status       RADIAN_READ_OK, RADIAN_READ_EMPTY_SIGNAL or RADIAN_READ_MAD_ZERO per read (the
 *                   reference raises ValueError and basecall.py:79-82 skips the read); the call
 *                   itself still returns RADIAN_OK.
 */
int radian_normalise_batch_host(const int16_t *signal, const int64_t *offsets, int n_reads,
                                double outlier_z_score, int outlier_is_int, void *out,
                                int32_t *out_is_int64, int32_t *out_status, int device);
/* the same on device pointers (offsets relative to `signal` / `out`), asynchronous on `stream` */
int radian_normalise_batch_dev(const int16_t *signal, const int64_t *offsets, int n_reads,
                               double outlier_z_score, int outlier_is_int, void *out,
                               int32_t *out_is_int64, int32_t *out_status, radian_stream_t stream);

/*
 * Windowing: replaces preprocess.get_windows (radian/preprocess.py:4-21, call site
 * radian/basecall.py:83).  radian_windows_plan gives, per read, the number of windows (the full
 * ones plus the zero-padded last one) and pad_end; radian_windows_batch_host writes the windows of
 * read r, `window` values each, at out[window_offsets[r] * window ...].  Values are copied as
 * 8-byte words, so float64 and int64 signals both work.  RADIAN_E_ARG for step <= 0 or
 * step > window (ValueError in the reference).
 */
int radian_windows_plan(const int64_t *offsets, int n_reads, int window, int step, int64_t *n_windows,
                        int32_t *pad_end);
int radian_windows_batch_host(const double *norm, const int64_t *offsets, int n_reads, int window, int step,
                              const int64_t *window_offsets, double *out, int device);
int radian_windows_batch_dev(const double *norm, const int64_t *offsets, const int64_t *window_offsets,
                             int n_reads, int window, int step, double *out, radian_stream_t stream);

/*
 * FASTA records of a decoded batch as the reference writes them (radian/basecall.py:129):
 * ">{read_id}\n{sequence reversed to 5'->3'}\n" per read.  seq / seq_offsets / len are the outputs of
 * a radian_decode_batch_* or radian_stitch_batch_* call (host copies); ids holds all read ids back to
 * back, id_offsets[n_reads+1] delimits them; bases: the 4 output letters ("ACGT").  Record r is
 * written at out + out_offsets[r] and must have exactly 1 + id length + 1 + len[r] + 1 bytes there
 * (out_offsets[n_reads+1] is a running sum of those sizes).  Host only, no device needed.
 */
int radian_fasta_records_host(const uint8_t *seq, const int64_t *seq_offsets, const int64_t *len,
                              int n_reads, const char *ids, const int64_t *id_offsets, const char *bases,
                              char *out, const int64_t *out_offsets);

#ifdef __cplusplus
}
#endif
#endif /* RADIAN_B200_H */
