/*
 * trx.h -- C ABI of the B200-native exact flat-index top-k engine (libtrx.so).
 *
 * This is the drop-in boundary for TextReact's retrieval hot path.  Every entry point
 * replaces one call the reference makes into the third-party `faiss` object
 * (reference = thomas0809/textreact; file:line relative to its root):
 *
 *   trx_create      <- faiss.IndexFlatL2(d)          retrieve/retrieve_faiss.py:65
 *                      (faiss.IndexFlatIP(d) for the 768-d neural retriever, README.md:44-47)
 *   trx_add         <- index.add(train_fps)           retrieve/retrieve_faiss.py:66
 *   trx_search      <- index.search(query_fps, k)     retrieve/retrieve_faiss.py:71
 *   trx_set_groups  <- gold-removed mode, lifted from the consumer-side filter
 *                      `skip_gold_neighbor`           textreact/dataset.py:74-76
 *   trx_set_row_attr <- `--before` year restriction         retrieve/retrieve_faiss.py:102-103
 *   trx_search_self <- train->train search (queries are the corpus)  retrieve/retrieve_faiss.py:114-115
 *   trx_reset / trx_destroy <- lifetime of the index object built per split
 *                                                     retrieve/retrieve_faiss.py:62-74
 *   trx_merge_topk, trx_exchange_* <- (no reference counterpart) k-way merge of per-shard results for
 *                      the row-sharded multi-GPU mode: after an NCCL all-gather,
 *                      or fused with the gather over NVLink peer memory
 *
 * Conventions
 *   - plain C, no C++ types, no exceptions, never abort(): every call returns an int
 *     status (TRX_OK == 0) and leaves a message retrievable with trx_last_error()
 *     (thread-local).
 *   - `x`, `xq`, `excl`, `g`, `D`, `I` may each be HOST or DEVICE pointers (detected with
 *     cudaPointerGetAttributes).  Every call is complete on return: trx_search waits for
 *     its last batch (it has to read the count of queries that need the exact fallback).
 *     Work is ordered on `stream` (NULL: the index's own stream); inside one call the
 *     batches are software-pipelined (upload and launch of batch i+1 overlap batch i).
 *   - the caller owns every buffer it passes; the index owns its device copies
 *     (fp32 corpus, bf16 corpus, norms, groups, workspaces) until reset/destroy.
 *   - results: best first (IP: larger score; L2: smaller squared distance), ties by
 *     ascending id, ids are add-order row numbers (+ id offset), unfilled slots
 *     I = -1, D = -FLT_MAX (IP) / +FLT_MAX (L2).
 *   - there is NO CPU fallback: without a CUDA device of compute capability 10.x
 *     trx_create fails with TRX_ENODEV.
 */
#ifndef TRX_H_
#define TRX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRX_OK 0
#define TRX_EINVAL 1  /* bad shape / argument */
#define TRX_ENOMEM 2  /* device (or pinned host) allocation failed */
#define TRX_ECUDA 3   /* CUDA runtime / driver error (message has the string) */
#define TRX_ENODEV 4  /* no usable sm_100 device */

#define TRX_METRIC_INNER_PRODUCT 0 /* == faiss.METRIC_INNER_PRODUCT */
#define TRX_METRIC_L2 1            /* == faiss.METRIC_L2 */

/* trx_set_option("path", v) values */
#define TRX_PATH_AUTO 0   /* pick by batch size (default) */
#define TRX_PATH_EXACT 1  /* fp32 streaming scan + exact select (the certified fallback) */
#define TRX_PATH_STREAM 2 /* bf16 CUDA-core streaming prefilter + fp32 rescore */
#define TRX_PATH_UMMA 3   /* bf16 tcgen05/TMEM prefilter + fp32 rescore */

typedef struct trx_index trx_index;

typedef struct trx_stats_t {
    int64_t searches;          /* trx_search calls */
    int64_t queries;           /* query rows processed */
    int64_t queries_exact;     /* rows answered by the exact fp32 scan (fallback or forced) */
    int64_t queries_uncert;    /* rows whose bf16 prefilter failed the exactness certificate */
    int64_t queries_overflow;  /* rows whose candidate list overflowed */
    int64_t rescored;          /* candidate rows rescored in fp32 */
    int64_t candidates;        /* candidates emitted by the prefilter */
    int32_t last_path;         /* TRX_PATH_* taken by the last batch */
    int32_t sm_count;
    int64_t launches;          /* kernels of ours launched so far */
    double last_prefilter_ms;  /* device time of the dominant scoring kernel, last timed batch */
    double last_total_ms;      /* device time of the whole last batch (same condition) */
    /* Sums over every batch that carried event records (all batches except the small ones replayed as a CUDA
     * graph): differences of two trx_stats calls give the per-stage device time of the batches in between. */
    int64_t timed_batches;
    double sum_sample_ms;      /* batch begin + sample pass + thresholds */
    double sum_prefilter_ms;   /* main scoring pass (K2 tcgen05 / K3 streaming) incl. candidate scatter */
    double sum_rescore_ms;     /* K4 exact rescore + certificate + final top-k (batches without fallbacks) */
    double sum_total_ms;       /* whole batch up to the results being ready on the device */
    int64_t queries_second_pass; /* rows without a certificate answered by the batched second prefilter pass
                                    (threshold = k-th exact score seen - eps: a complete candidate list) */
} trx_stats_t;

/* Create an empty flat index of dimension d on CUDA device `device`. */
int trx_create(int d, int metric, int device, trx_index** out);

/* Append n rows of d floats (row-major, contiguous). */
int trx_add(trx_index* idx, const float* x, int64_t n);

/* Append n rows of d elements of type `dtype` (TRX_DTYPE_*), widened to fp32 on the device exactly as
 * the FAISS python wrapper's np.ascontiguousarray(x, dtype='float32') would on the host: the reference
 * passes int64 difference fingerprints and int8 Morgan bits (retrieve/retrieve_faiss.py:26, :40). */
#define TRX_DTYPE_F32 0
#define TRX_DTYPE_F64 1
#define TRX_DTYPE_F16 2
#define TRX_DTYPE_I8 3
#define TRX_DTYPE_U8 4
#define TRX_DTYPE_I16 5
#define TRX_DTYPE_I32 6
#define TRX_DTYPE_I64 7
int trx_add_typed(trx_index* idx, const void* x, int64_t n, int dtype);

/* Pre-size the device buffers for `n` total rows (optional; avoids regrowth copies). */
int trx_reserve(trx_index* idx, int64_t n);

/* Per-row exclusion group (text-dedup group / patent id), n must equal ntotal. */
int trx_set_groups(trx_index* idx, const int32_t* g, int64_t n);

/* Per-row integer attribute (e.g. the publication year: the reference restricts the corpus with
 * `train_df[train_df['year'] < args.before]`, retrieve/retrieve_faiss.py:102-103, a dataframe filter before
 * add).  With trx_set_option("attr_below", T) rows whose attribute is >= T are ineligible, so one resident
 * corpus serves every --before split (retrieve/retro_year.sh:12).  n must equal ntotal; NULL clears. */
int trx_set_row_attr(trx_index* idx, const int32_t* attr, int64_t n);

/* k nearest rows for each of nq queries.  excl (nullable): per-query group to exclude,
 * -1 = none; requires trx_set_groups.  D: float[nq*k], I: int64[nq*k]. */
int trx_search(trx_index* idx, const float* xq, int64_t nq, int k, const int32_t* excl,
               float* D, int64_t* I, void* cuda_stream);

/* trx_search with per-call modes (nothing is left set on the index).  NULL params == trx_search without mask. */
typedef struct trx_search_params_t {
    const int32_t* exclude;  /* nullable: per-query group to exclude (-1 = none), host or device */
    int32_t attr_below;      /* rows with attribute >= this are ineligible; 2147483647 = no filter */
    int32_t dedup_groups;    /* 1: distinct-groups mode (see "dedup_groups" below) */
    int64_t self_row0;       /* >= 0: the queries are the stored rows [self_row0, self_row0 + nq), xq is ignored */
} trx_search_params_t;
int trx_search_ex(trx_index* idx, const float* xq, int64_t nq, int k, const trx_search_params_t* params,
                  float* D, int64_t* I, void* cuda_stream);

/* The search of ONE batch (nq <= max_batch) in two halves, for the row-sharded multi-GPU mode (SURVEY.md section 8e;
 * no reference counterpart).  A shard that answers alone must rescore enough candidates to certify ITS top-k; but most
 * of a shard's top-k never reaches the global top-k.  So:
 *   trx_search_begin   prefilter up to the masked, sorted candidate lists; payload[q] (device, [nq, nb + 1] floats) =
 *                      the nb best prefilter scores of query q on this shard, then the query's certificate slack eps
 *   (the shards exchange the payloads: trx_exchange_floor -> floor[nq])
 *   trx_search_finish  exact rescore of the candidates at or above floor[q] only, then the usual results -- with
 *                      possibly fewer than k rows per query (padded with -1): exactly this shard's members of the
 *                      global top-k and a few more.  floor NULL: the plain local top-k.
 * The k-way merge of the shards' results is the exact global top-k.  The queries / mask are copied by begin; nothing
 * else may be searched on the index between the two calls. */
int trx_search_begin(trx_index* idx, const float* xq, int64_t nq, int k, const trx_search_params_t* params, int nb,
                     float* payload, void* cuda_stream);
int trx_search_finish(trx_index* idx, const float* floor, float* D, int64_t* I, void* cuda_stream);

/* trx_search with the stored rows [row0, row0+nq) as the queries -- the reference's train->train search
 * (query_fps is train_fps, retrieve/retrieve_faiss.py:114-115) without sending the corpus to the device twice. */
int trx_search_self(trx_index* idx, int64_t row0, int64_t nq, int k, const int32_t* excl,
                    float* D, int64_t* I, void* cuda_stream);

/* The stored fp32 rows [row0, row0+n) -> out (host or device): FAISS's index.reconstruct / reconstruct_n. */
int trx_reconstruct(trx_index* idx, int64_t row0, int64_t n, float* out);

/* Remove all rows (keeps d / metric / options). */
int trx_reset(trx_index* idx);
void trx_destroy(trx_index* idx);

int64_t trx_ntotal(const trx_index* idx);
int trx_dim(const trx_index* idx);
int trx_metric(const trx_index* idx);

/* Added to every returned id: the first global row of this shard (row-sharded mode). */
int trx_set_id_offset(trx_index* idx, int64_t offset);

/* Tunables: "path" (TRX_PATH_*), "max_batch", "target_candidates", "sample_rate",
 * "stream_max_batch" (crossover at or below which AUTO uses the streaming kernel),
 * "umma_pair" / "pair_min_batch" (CTA-pair tiling from this batch size on),
 * "pipeline" (0: serial batches), "thr_margin" (0: candidate thresholds exactly as sampled; default 1 keeps each at
 * least 2.5 eps under the sample's estimate of the query's k-th score, the margin the certificate needs), "second_pass" (0: uncertified queries take an fp32 streaming sweep per 4 queries
 * instead of one batched second tcgen05 pass), "attr_below" (see trx_set_row_attr; 2147483647 = off),
 * "dedup_groups" (1: distinct-groups mode -- of the rows that share a group (trx_set_groups) only the best
 * one is returned, so the k results are k different texts: the consumer's deduplicate_neighbors,
 * textreact/dataset.py:46-56 and :77, applied before truncation instead of after), "timing". */
int trx_set_option(trx_index* idx, const char* key, double value);
int trx_get_option(const trx_index* idx, const char* key, double* value);

int trx_stats(const trx_index* idx, trx_stats_t* out);

/* K-way merge of G per-shard results (each [nq,k], best first, global ids) into one.
 * Dg: float[G*nq*k], Ig: int64[G*nq*k], shard-major.  Device pointers, stream-ordered. */
int trx_merge_topk(int metric, const float* Dg, const int64_t* Ig, int G, int64_t nq, int k,
                   float* D, int64_t* I, void* cuda_stream);

/* The exchange step of the row-sharded mode as one kernel over NVLink peer memory (one process per GPU, one node):
 * every rank exports a buffer through CUDA IPC, maps the buffers of its peers, and the all-gather of the per-shard
 * lists is fused into the merge kernel (peer loads; a flag protocol over peer stores replaces NCCL).
 *   create  -> allocate this rank's export buffer for up to max_entries = nq*k results per exchange
 *   handle  -> its 64-byte CUDA IPC handle (exchange the handles of all ranks by any means, in rank order)
 *   connect -> map the peers (handles: world*64 bytes)
 *   merge   -> collective: every rank passes its local [nq,k] lists (device pointers, best first, global ids) and
 *              receives the merged top-k; stream-ordered, no host synchronisation. */
typedef struct trx_exchange trx_exchange;
int trx_exchange_create(int device, int rank, int world, int64_t max_entries, trx_exchange** out);
int trx_exchange_handle(trx_exchange* ex, unsigned char* handle64);
int trx_exchange_connect(trx_exchange* ex, const unsigned char* handles);
int trx_exchange_merge(trx_exchange* ex, int metric, const float* D_local, const int64_t* I_local, int64_t nq, int k,
                       float* D, int64_t* I, void* cuda_stream);
/* The same exchange with the merge itself sharded: this rank receives the merged top-k of queries [q0, q0 + nq_out)
 * only (D, I: [nq_out, k]) and loads 1/world of the peers' entries.  Every rank still passes all nq local lists, and
 * every rank must take part in every exchange (nq_out may be 0). */
int trx_exchange_merge_slice(trx_exchange* ex, int metric, const float* D_local, const int64_t* I_local, int64_t nq, int k,
                             int64_t q0, int64_t nq_out, float* D, int64_t* I, void* cuda_stream);
/* Bounds exchange of the two-phase search (trx_search_begin / trx_search_finish): every rank passes the payload its
 * trx_search_begin produced ([nq, nb + 1] floats: the nb best prefilter scores of every query on this shard, then the
 * query's certificate slack eps) and receives floor[nq]: a prefilter score below which no row of the GLOBAL top-k can
 * lie (k-th largest of the world * nb exchanged scores - 2 * the largest eps).  Collective, stream-ordered. */
int trx_exchange_floor(trx_exchange* ex, const float* payload, int64_t nq, int nb, int k, float* floor_out,
                       void* cuda_stream);
void trx_exchange_destroy(trx_exchange* ex);

/* Plain device buffers for callers without a CUDA runtime of their own (a C program, a numpy-only script): the
 * device-pointer forms of the calls above (trx_search_begin / _finish, trx_merge_topk, ...) can then be used without
 * torch.  trx_device_copy moves bytes host <-> device in any direction; it first waits for all work queued on the
 * current device (the index's streams are non-blocking) and is complete on return. */
#include <stddef.h>
int trx_device_malloc(int device, size_t bytes, void** out);
int trx_device_free(void* p);
int trx_device_copy(void* dst, const void* src, size_t bytes);

/* Raw bf16 scoring GEMM on the tcgen05 path, for tests and profiling:
 * out[nq, n] = bf16(xq) . bf16(x_row)  (fp32 accumulate) over rows [row0, row0+n). */
int trx_debug_scores_umma(trx_index* idx, const float* xq, int64_t nq, int64_t row0, int64_t n,
                          float* out, void* cuda_stream);

const char* trx_last_error(void);
const char* trx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TRX_H_ */
