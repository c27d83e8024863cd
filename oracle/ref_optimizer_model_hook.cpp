/* TEST INFRASTRUCTURE (oracle/_ref/ref_host_optimizer_model.so, CPU only).  An extern "C" door to the reference's OWN CPU
 * model of the sparse optimizers -- class CPUOptimizer of its gradient-apply test
 * (cpp/tests/wholememory_ops/wholememory_embedding_gradient_apply_tests.cu:169-371) -- which its GPU test compares the
 * kernels against at 1e-5.  oracle/build_ref_host_optimizer_model.sh cuts the parameter struct and that class out of the
 * reference file at build time (by their opening / closing lines, into a temporary file that is deleted afterwards; no
 * reference source is kept) and compiles them together with this file.  tests/test_ref_optimizer_model.py checks this
 * repo's oracle (dedup + optimizer restatement) against it. */
#include <gtest/gtest.h>

#include <wholememory/embedding.h>

#include <cmath>
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include WGREF_OPTIMIZER_MODEL_SLICE /* struct EmbeddingBackwardTestParams ... class CPUOptimizer, verbatim from the reference tree */

extern "C" int wgref_cpu_optimizer_run(int optimizer_type,
                                       int64_t rows,
                                       int dim,
                                       const char* const* param_names,
                                       const float* param_values,
                                       int param_count,
                                       float lr,
                                       int steps,
                                       const int64_t* step_counts,
                                       const int64_t* indices, /* all steps, concatenated */
                                       const float* grads,     /* all steps, concatenated, [n, dim] each */
                                       float* table /* [rows, dim], updated in place */)
{
  EmbeddingBackwardTestParams params;
  params.set_entry_count(rows).set_embedding_dim(dim);
  params.optimizer_type = static_cast<wholememory_optimizer_type_t>(optimizer_type);
  for (int i = 0; i < param_count; ++i) params.optimizer_params[param_names[i]] = param_values[i];
  std::vector<std::vector<float>> embs(rows, std::vector<float>(dim));
  for (int64_t r = 0; r < rows; ++r)
    for (int c = 0; c < dim; ++c) embs[r][c] = table[r * dim + c];
  CPUOptimizer cpu_optimizer(&params, 0, rows);
  int64_t pos = 0;
  for (int s = 0; s < steps; ++s) {
    /* duplicate gradients are merged exactly as the reference test does before calling Apply (same file, :440-465):
     * first occurrence fixes the slot, later ones are added in arrival order */
    std::vector<int64_t> uniq;
    std::vector<std::vector<float>> merged;
    std::unordered_map<int64_t, int> slot;
    for (int64_t i = 0; i < step_counts[s]; ++i, ++pos) {
      const int64_t idx = indices[pos];
      const float* g    = grads + pos * dim;
      auto it           = slot.find(idx);
      if (it == slot.end()) {
        slot[idx] = (int)uniq.size();
        uniq.push_back(idx);
        merged.emplace_back(g, g + dim);
      } else {
        for (int d = 0; d < dim; ++d) merged[it->second][d] += g[d];
      }
    }
    cpu_optimizer.Apply(lr, uniq, merged, embs);
  }
  for (int64_t r = 0; r < rows; ++r)
    for (int c = 0; c < dim; ++c) table[r * dim + c] = embs[r][c];
  return 0;
}
