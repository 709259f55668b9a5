/* C restatement of DiffNorm's run-length unit reduction.
 *
 * TEST INFRASTRUCTURE ONLY (oracle): never linked into or called by the product library.
 * Follows research/TranSpeech/diff_norm_synthesis.py:25-46 (reduce_token), which is identical to
 * fairseq/data/audio/repr_to_repr_unit_dataset.py:92-113 (_reduce_tgt).
 *
 * tokens[0..n) -> dedup[0..R), duration[0..R') , index_to_keep[0..R); returns R.
 * Quirk kept from the reference: the trailing duration is appended unconditionally, so n == 0 yields
 * R = 0 runs but ONE duration entry (value 1); *n_dur receives the number of duration entries.
 */
#include <stdint.h>

int64_t dn_oracle_reduce_tgt(const int64_t *tokens, int64_t n, int64_t *dedup, int64_t *duration,
                             int64_t *index_to_keep, int64_t *n_dur)
{
    int64_t r = 0, d = 0, acc = 1;
    for (int64_t i = 0; i < n; ++i) {
        if (i == 0) {
            dedup[r] = tokens[i];
            index_to_keep[r] = i;
            ++r;
        } else if (tokens[i] != tokens[i - 1]) {
            duration[d++] = acc;
            dedup[r] = tokens[i];
            index_to_keep[r] = i;
            ++r;
            acc = 1;
        } else {
            ++acc;
        }
    }
    duration[d++] = acc;
    *n_dur = d;
    return r;
}

/* batched helper for the CPU baseline: B utterances packed row-major [B, T] with lengths[b] valid tokens.
 * Outputs are packed per row into [B, T] buffers; counts[b] = number of runs. */
void dn_oracle_reduce_tgt_batch(const int64_t *tokens, const int64_t *lengths, int64_t B, int64_t T,
                                int64_t *dedup, int64_t *duration, int64_t *index_to_keep, int64_t *counts)
{
    for (int64_t b = 0; b < B; ++b) {
        int64_t nd;
        /* duration needs room for the n == 0 quirk entry: callers give T >= 1 */
        counts[b] = dn_oracle_reduce_tgt(tokens + b * T, lengths[b], dedup + b * T, duration + b * T,
                                         index_to_keep + b * T, &nd);
    }
}
