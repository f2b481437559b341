/*
 * pgwin.h — C-ABI of libpgwin.so: the B200 (sm_100a) engine for the per-window / per-site numerics of
 * simonhmartin/genomics_general (popgenWindows.py / ABBABABAwindows.py / fourPopWindows.py / freq.py / sfs.py /
 * distMat.py) and for the .geno text parsing in front of them.
 *
 * The reference has no FFI: its seam is the Python API of genomics.py as used by the four scripts
 * (SURVEY.md §8b).  Each entry point below names the reference code it replaces
 * (/root/reference/<file>:<line>).  The binding a maintainer would add is a ctypes stub —
 * see INTEGRATION.md and genomics_general_b200/_lib.py.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; pg_last_error() gives the message of the
 *     last failure on the calling thread.
 *   - the caller owns every host buffer; the library owns device memory behind the opaque pg_ctx.
 *   - one ctx per device; calls on one ctx are not thread-safe; different ctxs are independent.
 *   - genotype codes: A=0 C=1 G=2 T=3, missing = any value with bit 7 set (canonically -1).
 *     This is Alignment.numArray (genomics.py:74-77, 834) narrowed to int8 and transposed to
 *     site-major: geno[site * H + hap].
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef PGWIN_H
#define PGWIN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pg_ctx pg_ctx;

/* ---- library / context ------------------------------------------------------------------------ */
int         pg_version(void);
const char* pg_last_error(void);
int         pg_device_count(int* n);
int         pg_ctx_create(int device, pg_ctx** out);
int         pg_ctx_destroy(pg_ctx* ctx);
/* pinned host memory for the end-to-end path (H2D from pinned buffers) */
int         pg_host_alloc(void** ptr, size_t bytes);
int         pg_host_free(void* ptr);

/* ---- data in ---------------------------------------------------------------------------------- */
/* Replaces genoToAlignment + Alignment.__init__ (genomics.py:1101-1127, 813-869): the dense matrix of
 * a whole file (or shard).  geno: int8 [S x H] site-major, pos: int32 [S] (may be NULL -> zeros).
 * Copies host->device (pitched so that rows are 16-byte multiples). */
int pg_upload(pg_ctx* ctx, const int8_t* geno, int64_t S, int32_t H, const int32_t* pos);
/* Two-step variant: allocate, then fill site ranges (each call is an async H2D on the ctx stream). */
int pg_alloc_sites(pg_ctx* ctx, int64_t S, int32_t H);
int pg_upload_range(pg_ctx* ctx, int64_t site0, int64_t n, const int8_t* geno, const int32_t* pos);
/* Appends n sites (int8 [n x H], reference codes; pos int32 [n] or NULL) after the resident ones; the matrix grows.  Used by
 * the multi-GPU command lines for the halo sites of windows that reach into the next rank's share of the file. */
int pg_append_sites(pg_ctx* ctx, int64_t n, const int8_t* geno, const int32_t* pos);
/* Device-side synthetic data (SURVEY.md §8d distribution; bit-identical to synth.py::synth_genotypes /
 * synth_positions).  Thresholds are 32-bit integer probabilities (p * 2^32). */
int pg_synth_fill(pg_ctx* ctx, int64_t S, int32_t n_pops, int32_t samples_per_pop, int32_t ploidy,
                  uint64_t seed, uint64_t thr_var, uint64_t thr_out0, uint64_t thr_third, uint64_t thr_miss,
                  int32_t spacing);
/* Device -> host copy of a site range (tests / building host buffers for the e2e bench). */
int pg_download(pg_ctx* ctx, int64_t site0, int64_t n, int8_t* geno, int32_t* pos);

/* Replaces SampleData/Alignment.groups (genomics.py:1264-1290, 857-861): hap_pop[h] in [0,P) or -1. */
int pg_set_pops(pg_ctx* ctx, int32_t P, const int32_t* hap_pop);
/* Replaces the window generators' output (genomics.py:1971-2171): half-open site-index ranges
 * [lo[w], hi[w]) computed on the host with the generators' exact semantics (windows.py). Ranges may
 * overlap and may be empty. */
int pg_set_windows(pg_ctx* ctx, int64_t W, const int64_t* lo, const int64_t* hi);

/* ---- statistics ------------------------------------------------------------------------------- */
/* Replaces Alignment.groupDistStats (genomics.py:956-995) + the sites/mid bookkeeping of
 * popgenWindows.py:37-41 for every window.
 *   pi  [W x P], dxy/fst [W x P(P-1)/2] (pairs in (0,1),(0,2),..,(1,2).. order), n_sites [W],
 *   pos_sum [W] (sum of positions, for midPos genomics.py:1795), path [W]: 0 = failed (sites < minSites),
 *   1 = closed-form allele-count path (K1), 2 = pairwise path (K2).
 * force_path: 0 = route per window (K1 when exact, else K2), 2 = pairwise for every window.
 * min_sites <= 0 disables the n_ij < minSites mask (genomics.py:958). */
int pg_popgen(pg_ctx* ctx, int32_t min_sites, double min_data, int32_t force_path,
              double* pi, double* dxy, double* fst, int64_t* n_sites, int64_t* pos_sum, int32_t* path);

/* Same statistics, left on the DEVICE as fixed-width records (so that the multi-GPU all-gather can read
 * them in place): d_rec is a device buffer of W * (4 + 5P + 2*npairs) 8-byte words per window,
 *   [sites (int64), pos_sum (int64), path (int64), pi[P], dxy[npairs], fst[npairs],
 *    l, S[P], thetaPi[P], thetaW[P], TajD[P]] (statistics are doubles).
 * *n_pairwise (optional) receives the number of windows that went through the pairwise path. */
int pg_popgen_device(pg_ctx* ctx, int32_t min_sites, double min_data, int32_t force_path, void* d_rec,
                     int64_t* n_pairwise);

/* Replaces Alignment.groupFreqStats (genomics.py:1002-1028; popgenWindows --analysis popFreq): the columns
 * computed alongside by the most recent pg_popgen call on this ctx (after pg_set_freqstats(ctx, 1)).  l [W] = sites complete in every haplotype
 * that belongs to a population; S, theta_pi, theta_w, taj_d [W x P]. */
int pg_popgen_freqstats(pg_ctx* ctx, double* l, double* S, double* theta_pi, double* theta_w, double* taj_d);
/* The popFreq counters cost ~4 % of the site pass, so they are opt-in: enable before pg_popgen. */
int pg_set_freqstats(pg_ctx* ctx, int32_t enable);

/* Replaces genomics.ABBABABA (genomics.py:1647-1695, polarize=True) per window.
 * out [W x 5] = ABBA, BABA, D, fd, fdM; sites_used [W] (double: nan when the window has no good site,
 * genomics.py:1694-1695); n_sites/pos_sum as above. p1,p2,p3,o are population indices of pg_set_pops. */
int pg_abbababa(pg_ctx* ctx, int32_t p1, int32_t p2, int32_t p3, int32_t o, double min_data,
                double* out, double* sites_used, int64_t* n_sites, int64_t* pos_sum);

/* Replaces genomics.fourPop (genomics.py:1585-1643; fourPopWindows.py:27-52) per window.
 * out [W x 14] = fhom, fhom', D, fd, fd', fdm, fdm', fdh, fdh2, fh, ABBA, BABA, ABAA, BAAA; sites_used [W] (0 when the
 * window has no good site, 1641-1643).  mode 0 = default (the allele np.argsort(all4freqs)[:,2] picks, i.e. the rarer of
 * the two; on an exact tie the reference depends on numpy's sort implementation — here the lower allele index),
 * 1 = polarize (allele absent from P4), 2 = fixed (polarize + P1,P2,P3 each fixed). */
int pg_fourpop(pg_ctx* ctx, int32_t p1, int32_t p2, int32_t p3, int32_t p4, double min_data, int32_t mode,
               double* out, double* sites_used, int64_t* n_sites, int64_t* pos_sum);

/* Replaces Alignment.siteFreqs(asCounts=True) per population (genomics.py:1049-1052; freq.py:52-58):
 * counts uint16 [n x P x 4] (A,C,G,T) for sites site0 .. site0+n-1. */
int pg_site_counts(pg_ctx* ctx, int64_t site0, int64_t n, uint16_t* counts);

/* Replaces freq.py --target derived|minor (freq.py:62-92; derivedAllele / minorAllele genomics.py:636-668):
 * out double [n x P] = frequency (count/non-missing; nan = no value) or, with as_counts, count (0 = no value) of the
 * target allele in each population.  target 1 = derived (the LAST population is the outgroup), 2 = minor allele over
 * all populations' haplotypes.  min_data is compared with the population's non-missing COUNT, as the reference does
 * (freq.py:79).  The reference draws at random when the two alleles are exactly tied (genomics.py:667): here the
 * lower allele is used and tie[s] = 1 (tie may be NULL).  Values are not rounded (freq.py:91 rounds to 4 dp). */
int pg_site_target_freqs(pg_ctx* ctx, int64_t site0, int64_t n, int32_t target, double min_data, int32_t as_counts,
                         double* out, uint8_t* tie);

/* Replaces the per-site loop of sfs.py for --inputType genotypes without subsampling (sfs.py:430-470: population base
 * counts, the completeness test of 449, getTargetCounts 68-92) and the SparseFS accumulation (94-125, 484-487).
 * In-group = populations 0..n_in-1 of pg_set_pops; outgroup = a later population index (polarized spectra) or -1 (the
 * second most frequent allele, `totalBaseCounts.argsort()[-2]`; on an exact tie of the two alleles the reference depends
 * on numpy's sort, here the lower allele).  Spectrum g is over the populations group_pops[group_off[g] .. group_off[g+1]);
 * hist holds the spectra one after the other as dense row-major arrays of prod(N_k + 1) cells, first[cell] = index of
 * the first site that hit the cell (-1 = empty; gives the reference's first-appearance output order).
 * site_mask (may be NULL): only sites with mask 1 are counted (--include / --exclude). */
int pg_sfs(pg_ctx* ctx, int32_t n_in, int32_t outgroup, int32_t n_groups, const int32_t* group_off,
           const int32_t* group_pops, const uint8_t* site_mask, int64_t* hist, int64_t* first, int64_t* n_counted);

/* The same spectra from TABLES of counts on the host (sfs.py --inputType baseCounts | targetCounts, sfs.py:456-474):
 * kind 0: table = uint16 [n x P x 4] base counts per population (the rows freq.py writes); kind 1: int32 [n x P] counts of
 * the target allele (sfs.py's default input, e.g. freq.py --target derived --asCounts).  No completeness test.
 * dims[X] = radix of population X in the dense histograms (its largest count + 1). */
int pg_sfs_tables(pg_ctx* ctx, int32_t kind, const void* table, int64_t n, int32_t P, const int32_t* dims, int32_t n_in,
                  int32_t outgroup, int32_t n_groups, const int32_t* group_off, const int32_t* group_pops,
                  const uint8_t* site_mask, int64_t* hist, int64_t* first, int64_t* n_counted);

/* Replaces Alignment.indPairDists (genomics.py:934-954) as used by distMat.py:42-45 and popgenWindows.py:54-57.
 * hap_ind[h] = individual index in [0,n_ind) or -1; dist [W x n_ind x n_ind]; n_sites/pos_sum [W] (may be NULL).
 * min_sites > 0: haplotype pairs with n_ij < min_sites are nan — the state of the reference's cached matrix when
 * groupDistStats ran earlier on the same window (it masks in place, genomics.py:959-961); 0 = no mask (distMat.py). */
int pg_pairdist(pg_ctx* ctx, int32_t n_ind, const int32_t* hap_ind, int32_t include_same_with_same,
                int32_t min_sites, double* dist, int64_t* n_sites, int64_t* pos_sum);

/* Replaces distMat.py --windType cat (distMat.py:303-314: parseGenoFile turns the WHOLE file into one window, then
 * indPairDists): dist [n_ind x n_ind] over every uploaded site.  With a communicator (pg_nccl_init, nranks > 1) the
 * uploaded sites are this rank's shard of that window: the integer pair matrices diff_ij / n_ij are added across the
 * ranks with ONE ncclAllReduce (int64 sum) before the division, and every rank receives the same matrix.
 * *total_sites (may be NULL) = number of sites over all ranks. */
int pg_pairdist_cat(pg_ctx* ctx, int32_t n_ind, const int32_t* hap_ind, int32_t include_same_with_same,
                    double* dist, int64_t* total_sites);

/* Replaces Alignment.seqNonNan() (genomics.py:1038-1040) per window — the --minPerInd gate of distMat.py:40:
 * out int64 [W x H] = non-missing sites of each haplotype (upload order) inside each window. */
int pg_seq_nonnan(pg_ctx* ctx, int64_t* out);

/* Replaces Alignment.sampleHet() (genomics.py:918-929; popgenWindows.py:59-61 --analysis indHet): het [W x n_ind] =
 * p-distance between the two haplotypes of each individual; nan unless the individual has exactly two haplotypes
 * and bit 1 of n_ij is set (the reference's `len(x)==2 & n >= 1` is the chained comparison len(x) == (2 & n) >= 1).
 * min_sites as in pg_pairdist. */
int pg_ind_het(pg_ctx* ctx, int32_t n_ind, const int32_t* hap_ind, int32_t min_sites, double* het);

/* Replaces Alignment.H12stats(maxDist) + distMat_to_cluster_sizes (genomics.py:1079-1098, 1239-1261;
 * popgenWindows.py:63-64 --analysis hapStats): out [W x P x 3] = H1, H12, H2 for the populations of pg_set_pops.
 * min_sites as in pg_pairdist; diag_nan != 0 when an earlier groupDistStats / indPairDists of the same window set
 * the cached matrix's diagonal to nan (963, 940), which removes the self-matches from the greedy clustering. */
int pg_hapstats(pg_ctx* ctx, double max_dist, int32_t min_sites, int32_t diag_nan, double* out);

/* Replaces Alignment.distMatrix + pairNonNan (genomics.py:907-916, 1042-1047) for ONE window:
 * diff, n int32 [H x H] (symmetric, diagonal: diff 0, n = non-missing sites of the haplotype). */
int pg_pair_counts(pg_ctx* ctx, int64_t window, int32_t* diff, int32_t* n);

/* ---- multi-GPU: native NCCL all-gather of the per-window records -------------------------------- */
/* One process per GPU.  Windows shard across ranks with no data-path collective; the only exchange is ONE
 * ncclAllGather of the fixed-width records, enqueued on the ctx stream right behind the statistics kernels.
 * NCCL is bound at run time (dlopen): rank 0 creates the 128-byte id, the host distributes it by any means
 * (bench.py uses torch.distributed's broadcast), every rank calls pg_nccl_init.
 * pg_popgen_allgather: h_table (host) receives nranks * w_max * (4 + 5P + 2*npairs) 8-byte words in rank order
 * (record layout of pg_popgen_device; rows past a rank's own window count are zero); w_max >= every rank's W. */
int pg_nccl_unique_id(void* id128);
int pg_nccl_init(pg_ctx* ctx, int32_t nranks, int32_t rank, const void* id128);
int pg_nccl_finalize(pg_ctx* ctx);
int pg_popgen_allgather(pg_ctx* ctx, int32_t min_sites, double min_data, int32_t force_path, int64_t w_max,
                        void* h_table, int64_t* n_pairwise);
/* Pipelined form (two slots): `begin` enqueues site pass + finalize on the ctx stream and the exchange + read-back of the
 * table on a side stream; `end` waits for that batch and returns the slot's pinned table (nranks * w_max records, valid until
 * the slot's next `begin`).  begin(0); begin(1); end(0); begin(0); end(1); ... hides the all-gather and the D2H of one batch
 * under the site pass of the next (the reference's sorter + writer run concurrently with its workers,
 * popgenWindows.py:108-160).  Works without a communicator too (one rank). */
int pg_popgen_gather_begin(pg_ctx* ctx, int32_t min_sites, double min_data, int64_t w_max, int32_t slot);
int pg_popgen_gather_end(pg_ctx* ctx, int32_t slot, const void** h_table, int64_t* n_pairwise);
/* The same for the ABBA-BABA statistics (ABBABABAwindows.py window-sharded over the GPUs): h_table receives
 * nranks * w_max * 8 words per window [sites (int64), pos_sum (int64), ABBA, BABA, D, fd, fdM, sitesUsed (doubles)],
 * and for genomics.fourPop: 17 words [sites, pos_sum, the 14 statistics in pg_fourpop's order, sitesUsed]. */
int pg_abbababa_allgather(pg_ctx* ctx, int32_t p1, int32_t p2, int32_t p3, int32_t o, double min_data, int64_t w_max,
                          void* h_table);
int pg_fourpop_allgather(pg_ctx* ctx, int32_t p1, int32_t p2, int32_t p3, int32_t p4, double min_data, int32_t mode,
                         int64_t w_max, void* h_table);

/* ---- host-side .geno text ingest (no CUDA) ------------------------------------------------------ */
/* Replaces parseGenoLine/GenoFileReader (genomics.py:1884-1945) + splitSeq/haplo/forceHomo (390-396, 27, 407)
 * + seqArrayToNumArray (74-77) for a whole file: `buf` holds complete data lines (no header line);
 * '#' lines and blank lines are skipped.  fmt: 0 phased, 1 diplo, 2 pairs, 3 haplo.
 * col_take[k] = genotype column (0-based, after scaffold and position) of output sample k, ploidy[k] its
 * haplotype count; output row layout: samples in the given order, haplotypes of a sample adjacent.
 * geno [n_lines x H_out] int8, pos [n_lines] int32, new_scaffold [n_lines] (1 where the scaffold field differs
 * from the previous data line), line_off [n_lines] byte offset of each data line in buf. */
int pg_geno_count_lines(const char* buf, size_t len, int64_t* n_lines);
int pg_geno_parse(const char* buf, size_t len, int32_t fmt, int32_t n_out, const int32_t* col_take,
                  const int8_t* ploidy, int32_t H_out, int64_t n_lines, int8_t* geno, int32_t* pos,
                  int8_t* new_scaffold, int64_t* line_off, int32_t n_threads);

/* ---- host-side output rows (no CUDA) ------------------------------------------------------------- */
/* Replaces the row assembly of freq.py (freq.py:100-111) for n sites: "<scaffold>\t<position>\t<col>...\n".
 * mode 0: data = counts uint16 [n x P x 4], a column is "cA,cC,cG,cT"; mode 1: data = double [n x P] printed like
 * numpy's float64 -> str ("0.25", "1.0", "nan"); mode 2: double [n x P] printed as integers (--asCounts).
 * keep (may be NULL): uint8 [n], rows with 0 are skipped.  Thread t formats its share of the sites into
 * out + t * seg_cap and reports the bytes written in seg_len[t]; the caller writes the segments in order. */
int pg_format_freq_rows(int32_t mode, const void* data, int64_t n, int32_t P, const int32_t* pos, const int32_t* scaf_id,
                        const char* const* scaf_names, const uint8_t* keep, char* out, size_t seg_cap, int32_t n_threads,
                        size_t* seg_len);

/* Rows of a float64 matrix [rows x cols] as text, "<prefix[r]><v0><sep><v1>...\n" with numbers printed like numpy's
 * float64 -> str: the " ".join(row) over ndarray.round(roundTo).astype(str) of makeDistMat*String (genomics.py:2288-2306)
 * for distMat.py's per-window matrices (the caller rounds).  prefix may be NULL.  Segments as in pg_format_freq_rows. */
int pg_format_matrix_rows(const double* v, int64_t rows, int32_t cols, int32_t sep, const char* const* prefix, char* out,
                          size_t seg_cap, int32_t n_threads, size_t* seg_len);

/* ---- device-side .geno text ingest --------------------------------------------------------------- */
/* Same job and grammar as pg_geno_parse, on the GPU: the text (complete data lines, no header line) is copied to
 * device memory as it is and tokenised there, straight into this ctx's resident matrix (replaces pg_geno_parse +
 * pg_upload when the text fits in device memory).  col_hap[c] = first output haplotype of genotype column c
 * (0-based, after scaffold and position) or -1 for a column that is not wanted; col_ploidy[c] its haplotype count.
 * *n_sites = number of data lines = sites now resident.  Errors (a token whose allele count does not match the
 * ploidy, genomics.py:1111; a non-integer position; a short line) name the offending data line. */
int pg_ingest_text(pg_ctx* ctx, const char* buf, size_t len, int32_t fmt, int32_t n_cols, const int32_t* col_hap,
                   const int8_t* col_ploidy, int32_t H_out, int64_t* n_sites);
/* The same for a file on disk: bytes [body_offset, EOF) of `path` (body_offset = length of the header line, or 0) are
 * read straight into the pinned staging buffers by a few host threads — the file is never copied whole into host
 * memory.  line_off values of pg_ingest_meta are relative to body_offset. */
int pg_ingest_file(pg_ctx* ctx, const char* path, int64_t body_offset, int32_t fmt, int32_t n_cols, const int32_t* col_hap,
                   const int8_t* col_ploidy, int32_t H_out, int64_t* n_sites);
/* One rank's share of a file in the multi-GPU command lines (replaces the single producer that feeds the -T workers,
 * popgenWindows.py:398-446): bytes [byte_lo, byte_hi) — both at line starts, byte_hi < 0 = end of file.  line_off values of
 * pg_ingest_meta are relative to byte_lo. */
int pg_ingest_file_range(pg_ctx* ctx, const char* path, int64_t byte_lo, int64_t byte_hi, int32_t fmt, int32_t n_cols,
                         const int32_t* col_hap, const int8_t* col_ploidy, int32_t H_out, int64_t* n_sites);
/* pos int32 [S], new_scaffold int8 [S], line_off int64 [S] of the last pg_ingest_text / pg_ingest_file (as pg_geno_parse returns them;
 * any pointer may be NULL). */
int pg_ingest_meta(pg_ctx* ctx, int32_t* pos, int8_t* new_scaffold, int64_t* line_off);
/* Frees the device copy of the text. */
int pg_ingest_release(pg_ctx* ctx);

/* ---- introspection ---------------------------------------------------------------------------- */
/* Device time (ms, CUDA events on the ctx stream) of the kernels launched by the last statistics call:
 * names[i] -> ms[i]; returns the number of entries through *count (at most cap). */
int pg_last_timings(pg_ctx* ctx, int32_t cap, char (*names)[32], float* ms, int32_t* launches, int32_t* count);
/* Total kernels launched by this ctx so far. */
int pg_launch_count(pg_ctx* ctx, int64_t* n);
/* Host-only planning self-test hook (no device needed): returns the K1 launch plan for a shape. */
int pg_debug_k1_plan(int64_t S, int32_t H, int32_t* pitch, int32_t* lanes_per_site, int32_t* tile_sites,
                     int32_t* stages, int32_t* smem_bytes);

#ifdef __cplusplus
}
#endif
#endif /* PGWIN_H */
