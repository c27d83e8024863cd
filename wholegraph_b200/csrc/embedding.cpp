/*
 * Embedding objects + sparse optimizers behind the C API of include/wholememory/embedding.h.
 * Behaviour follows reference cpp/src/wholememory/embedding.cpp (allocate :87-144, gradient path
 * :146-323, optimizer states :325-428, C API :900-1152) and embedding_optimizer.cpp (parameters
 * :65-75, defaults/state names :170-190, :307, :404-410) for the NON-CACHED embedding.
 *
 * Gradient path here:  bucket ids by owner -> (id, fp32 gradient row) pairs reach their owners by peer stores over
 * NVLink (peer_push.cu; NCCL all-to-all in exchange.cu when the GPUs cannot map each other; nothing at all on a
 * 1-rank communicator) -> ONE fused merge+update kernel (sparse_optimizer.cu).
 * The reference runs: bucket/sort, gather-permute, alltoallv, sort, unique_by_key, dedup kernel,
 * optimizer kernel.
 *
 * Deliberate deviations (DESIGN.md "deviations"):
 *  - optimizer state tensors returned by wholememory_embedding_get_optimizer_state cover all N rows
 *    (the reference slices rows [0, D) by mistake: embedding.cpp:336 passes sizes[1] as the row end).
 *  - cache policies and round-robin sharding are refused (out of scope), not silently ignored.
 */
#include "exchange.hpp"
#include "sparse_optimizer.hpp"

#include <string>

struct wholememory_embedding_cache_policy_ {
  wholememory_comm_t cache_comm;
  wholememory_memory_type_t memory_type;
  wholememory_memory_location_t memory_location;
  wholememory_access_type_t access_type;
  float ratio;
};

struct wholememory_embedding_optimizer_ {
  wholememory_optimizer_type_t type = WHOLEMEMORY_OPT_NONE;
  wm::optimizer_params params;
  std::vector<const char*> state_names; /* nullptr-terminated */
};

struct wholememory_embedding_ {
  wholememory_comm_t comm           = nullptr;
  wholememory_tensor_t allocated    = nullptr; /* [N, padded stride] */
  wholememory_tensor_t user         = nullptr; /* [N, D] view */
  wholememory_dtype_t dtype         = WHOLEMEMORY_DT_UNKNOWN;
  int gather_sms                    = -1;
  wholememory_embedding_optimizer_t optimizer = nullptr;
  /* optimizer state */
  wholememory_embedding_t state_embedding = nullptr; /* [N, state_stride] fp32, same type/location/partition */
  std::vector<std::pair<std::string, wholememory_tensor_t>> named_states;
  wholememory_tensor_t b12_padded = nullptr, b12_user = nullptr;
  wm::push_stage grad_stage; /* peer-store gradient exchange (peer_push.cu), allocated on first use */
};

namespace wm {
namespace {

int64_t pad_to_16_bytes(int64_t dim, size_t esize)
{
  int64_t per = 16 / (int64_t)esize;
  return (dim + per - 1) / per * per; /* reference align_embedding_dim, embedding.cpp:43-50 */
}

void* local_rows_ptr(wholememory_tensor_t t, size_t* row_start, size_t* rows)
{
  void* p      = nullptr;
  size_t bytes = 0, off = 0;
  if (wholememory_get_local_memory(&p, &bytes, &off, wholememory_tensor_get_memory_handle(t)) != WHOLEMEMORY_SUCCESS)
    WM_THROW(WHOLEMEMORY_LOGIC_ERROR, "cannot resolve the local shard");
  WM_EXPECT(wholememory_tensor_get_local_entry_start(row_start, t) == WHOLEMEMORY_SUCCESS, WHOLEMEMORY_LOGIC_ERROR, "local entry start");
  WM_EXPECT(wholememory_tensor_get_local_entry_count(rows, t) == WHOLEMEMORY_SUCCESS, WHOLEMEMORY_LOGIC_ERROR, "local entry count");
  return p;
}

wholememory_error_code_t create_states(wholememory_embedding_t e)
{
  auto* opt       = e->optimizer;
  auto* udesc     = wholememory_tensor_get_tensor_description(e->user);
  auto* adesc     = wholememory_tensor_get_tensor_description(e->allocated);
  const int64_t N = udesc->sizes[0], D = udesc->sizes[1];
  const int64_t padded = adesc->strides[0];
  auto* h             = wholememory_tensor_get_memory_handle(e->allocated);
  const int ws        = e->comm->world_size;
  std::vector<size_t> part(ws);
  WHOLEMEMORY_RETURN_ON_FAIL(wholememory_tensor_get_entry_partition_sizes(part.data(), e->allocated));

  std::vector<std::string> elementwise; /* per-element states, each D wide, packed side by side */
  switch (opt->type) {
    case WHOLEMEMORY_OPT_LAZY_ADAM: elementwise = {"m", "v"}; break;
    case WHOLEMEMORY_OPT_ADAGRAD: elementwise = {"state_sum"}; break;
    case WHOLEMEMORY_OPT_RMSPROP: elementwise = {"v"}; break;
    default: break;
  }
  if (!elementwise.empty()) {
    wholememory_tensor_description_t sd = *udesc;
    sd.dtype                            = WHOLEMEMORY_DT_FLOAT;
    sd.sizes[1] = sd.strides[0] = padded * (int64_t)elementwise.size();
    sd.storage_offset           = 0;
    WHOLEMEMORY_RETURN_ON_FAIL(wholememory_create_embedding(&e->state_embedding, &sd, e->comm, wholememory_get_memory_type(h),
                                                            wholememory_get_memory_location(h), nullptr, part.data(), -1, 0));
    wholememory_tensor_t st = wholememory_embedding_get_embedding_tensor(e->state_embedding);
    for (size_t i = 0; i < elementwise.size(); ++i) {
      int64_t starts[2] = {0, (int64_t)i * padded};
      int64_t ends[2]   = {N, (int64_t)i * padded + D};
      wholememory_tensor_t sub = nullptr;
      WHOLEMEMORY_RETURN_ON_FAIL(wholememory_tensor_get_subtensor(st, starts, ends, &sub));
      e->named_states.emplace_back(elementwise[i], sub);
    }
    /* zero my shard (reference zero_local_state_tensor, embedding_optimizer.cpp:41-51) */
    size_t row0 = 0, rows = 0;
    void* p = local_rows_ptr(wholememory_embedding_get_embedding_tensor(e->state_embedding), &row0, &rows);
    if (rows > 0) {
      size_t bytes = rows * (size_t)sd.strides[0] * sizeof(float);
      if (wholememory_get_memory_location(h) == WHOLEMEMORY_ML_DEVICE) WM_CUDA(cudaMemset(p, 0, bytes));
      else memset(p, 0, bytes);
    }
  }
  if (opt->type == WHOLEMEMORY_OPT_LAZY_ADAM) {
    /* per-row beta1^t, beta2^t: [N, 2] fp32, DISTRIBUTED/DEVICE, same row partition (reference :407-425) */
    wholememory_tensor_description_t bd = *adesc;
    bd.dtype                            = WHOLEMEMORY_DT_FLOAT;
    bd.sizes[1] = bd.strides[0] = 2;
    WHOLEMEMORY_RETURN_ON_FAIL(wholememory_create_tensor(&e->b12_padded, &bd, e->comm, WHOLEMEMORY_MT_DISTRIBUTED, WHOLEMEMORY_ML_DEVICE, part.data()));
    int64_t starts[2] = {0, 0}, ends[2] = {N, 2};
    WHOLEMEMORY_RETURN_ON_FAIL(wholememory_tensor_get_subtensor(e->b12_padded, starts, ends, &e->b12_user));
    e->named_states.emplace_back("beta12t", e->b12_user);
    size_t row0 = 0, rows = 0;
    void* p = local_rows_ptr(e->b12_padded, &row0, &rows);
    fill_float(static_cast<float*>(p), 1.0f, (int64_t)rows * 2, nullptr);
  }
  WM_CUDA(cudaDeviceSynchronize());
  return WHOLEMEMORY_SUCCESS;
}

void destroy_states(wholememory_embedding_t e)
{
  for (auto& kv : e->named_states)
    if (kv.second != e->b12_user) wholememory_destroy_tensor(kv.second);
  e->named_states.clear();
  if (e->b12_user) wholememory_destroy_tensor(e->b12_user);
  if (e->b12_padded) wholememory_destroy_tensor(e->b12_padded);
  if (e->state_embedding) wholememory_destroy_embedding(e->state_embedding);
  e->b12_user = e->b12_padded = nullptr;
  e->state_embedding          = nullptr;
}

wholememory_error_code_t gradient_apply(wholememory_embedding_t e,
                                        wholememory_tensor_t indices,
                                        wholememory_tensor_t grads,
                                        float lr,
                                        wholememory_env_func_t* env,
                                        cudaStream_t stream)
{
  WM_EXPECT(e->optimizer != nullptr && e->optimizer->type != WHOLEMEMORY_OPT_NONE, WHOLEMEMORY_LOGIC_ERROR,
            "embedding has no optimizer: call wholememory_embedding_set_optimizer first");
  auto* idesc = wholememory_tensor_get_tensor_description(indices);
  auto* gdesc = wholememory_tensor_get_tensor_description(grads);
  auto* udesc = wholememory_tensor_get_tensor_description(e->user);
  auto* adesc = wholememory_tensor_get_tensor_description(e->allocated);
  WM_EXPECT(idesc->dim == 1 && (idesc->dtype == WHOLEMEMORY_DT_INT || idesc->dtype == WHOLEMEMORY_DT_INT64), WHOLEMEMORY_INVALID_INPUT,
            "indices must be a 1-D int32/int64 tensor");
  /* reference embedding_optimizer_func.cu:76-80 */
  WM_EXPECT(gdesc->dim == 2 && gdesc->dtype == WHOLEMEMORY_DT_FLOAT && gdesc->sizes[0] == idesc->sizes[0] && gdesc->strides[1] == 1,
            WHOLEMEMORY_INVALID_INPUT, "grads must be a [n, D] fp32 tensor with one row per index");
  WM_EXPECT(gdesc->sizes[1] == udesc->sizes[1], WHOLEMEMORY_INVALID_INPUT, "grads width %ld != embedding dim %ld",
            (long)gdesc->sizes[1], (long)udesc->sizes[1]);
  const int64_t n = idesc->sizes[0], D = udesc->sizes[1];
  const void* idx_ptr   = wholememory_tensor_get_data_pointer(indices);
  const float* grad_ptr = static_cast<const float*>(wholememory_tensor_get_data_pointer(grads));
  auto* comm            = e->comm;

  optimizer_rows rows{};
  size_t row0 = 0, nrows = 0;
  rows.w               = static_cast<float*>(local_rows_ptr(e->allocated, &row0, &nrows));
  rows.w_stride        = adesc->strides[0];
  rows.local_row_start = (int64_t)row0;
  rows.local_rows      = (int64_t)nrows;
  rows.dim             = (int)D;
  if (e->state_embedding) {
    size_t s0 = 0, sn = 0;
    wholememory_tensor_t st = e->state_embedding->allocated;
    rows.state              = static_cast<float*>(local_rows_ptr(st, &s0, &sn));
    rows.state_stride       = wholememory_tensor_get_tensor_description(st)->strides[0];
  }
  if (e->b12_padded) {
    size_t s0 = 0, sn = 0;
    rows.b12 = static_cast<float*>(local_rows_ptr(e->b12_padded, &s0, &sn));
  }
  const int64_t total_rows = udesc->sizes[0];

  if (comm->world_size == 1) {
    /* nothing to exchange: merge + update straight from the caller's buffers */
    merge_and_update_rows(e->optimizer->type, idx_ptr, idesc->dtype, n, grad_ptr, gdesc->strides[0], rows, e->optimizer->params, lr,
                          total_rows, /*may_have_negative=*/true, env, stream);
    return WHOLEMEMORY_SUCCESS;
  }

  auto* h = wholememory_tensor_get_memory_handle(e->allocated);
  const auto first_rows = handle_first_rows(h, (size_t)adesc->strides[0] * sizeof(float));
  exchange_plan plan(env);
  /* NVSwitch path: every rank stores its (id, gradient row) pairs straight into the owners' stages.  WG_GRAD_PUSH=0
   * keeps the NCCL all-to-all below (also used when the ranks' GPUs cannot map each other). */
  static const bool push_enabled = [] {
    const char* v = getenv("WG_GRAD_PUSH");
    return v == nullptr || v[0] != '0';
  }();
  if (push_enabled && comm->all_peer_capable && handle_is_addressable(h)) {
    partition_by_owner(&plan, comm, idx_ptr, idesc->dtype, n, first_rows, stream);
    const int64_t* recv_ids = nullptr;
    const float* recv_rows  = nullptr;
    int64_t n_recv = push_rows_to_owners(&e->grad_stage, comm, plan, grad_ptr, gdesc->strides[0], D, stream, &recv_ids, &recv_rows);
    merge_and_update_rows(e->optimizer->type, recv_ids, WHOLEMEMORY_DT_INT64, n_recv, recv_rows, D, rows, e->optimizer->params, lr,
                          total_rows, /*may_have_negative=*/false, env, stream);
    /* asynchronous from here: the stage is double-buffered and the plan's temporaries were consumed before the barrier */
    return WHOLEMEMORY_SUCCESS;
  }
  plan_exchange(&plan, comm, idx_ptr, idesc->dtype, n, first_rows, stream);
  temp_buffer outgoing(env), incoming(env);
  float* out_p = static_cast<float*>(outgoing.device((size_t)std::max<int64_t>(plan.n_send, 1) * D, WHOLEMEMORY_DT_FLOAT));
  float* in_p  = static_cast<float*>(incoming.device((size_t)std::max<int64_t>(plan.n_recv, 1) * D, WHOLEMEMORY_DT_FLOAT));
  if (plan.n_send > 0) {
    wholememory_matrix_description_t gm;
    wholememory_convert_tensor_desc_to_matrix(&gm, gdesc);
    gm.storage_offset = 0; /* get_data_pointer already applied it */
    int64_t sz[2]     = {plan.n_send, D};
    auto packed       = wholememory_create_matrix_desc(sz, D, 0, WHOLEMEMORY_DT_FLOAT);
    row_move(true, make_flat_table_ref(const_cast<float*>(grad_ptr)), gm, plan.origin.ptr(),
             wholememory_create_array_desc(plan.n_send, 0, WHOLEMEMORY_DT_INT64), out_p, packed, stream, -1);
  }
  exchange_rows(plan, comm, out_p, in_p, (size_t)D * sizeof(float), /*to_owner=*/true, stream);
  merge_and_update_rows(e->optimizer->type, plan.recv_idx.ptr(), idesc->dtype, plan.n_recv, in_p, D, rows, e->optimizer->params, lr,
                        total_rows, /*may_have_negative=*/false, env, stream);
  /* exchange buffers are released on return */
  WM_CUDA(cudaStreamSynchronize(stream));
  return WHOLEMEMORY_SUCCESS;
}

}  // namespace
}  // namespace wm

extern "C" {

wholememory_error_code_t wholememory_create_embedding_optimizer(wholememory_embedding_optimizer_t* optimizer,
                                                                wholememory_optimizer_type_t optimizer_type)
{
  if (optimizer == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  auto* o = new wholememory_embedding_optimizer_();
  o->type = optimizer_type;
  switch (optimizer_type) {
    case WHOLEMEMORY_OPT_SGD: o->state_names = {nullptr}; break;
    case WHOLEMEMORY_OPT_LAZY_ADAM: o->state_names = {"m", "v", "beta12t", nullptr}; break;
    case WHOLEMEMORY_OPT_ADAGRAD: o->state_names = {"state_sum", nullptr}; break;
    case WHOLEMEMORY_OPT_RMSPROP: o->state_names = {"v", nullptr}; break;
    default:
      delete o;
      WM_ERROR("unknown optimizer type %d", (int)optimizer_type);
      return WHOLEMEMORY_NOT_IMPLEMENTED; /* reference embedding_optimizer.cpp:527 */
  }
  wm::obj_register(wm::OBJ_OPTIMIZER, o);
  *optimizer = o;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_optimizer_set_parameter(wholememory_embedding_optimizer_t optimizer,
                                                             const char* parameter_name,
                                                             void* value)
{
  if (parameter_name == nullptr || value == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  WM_REQUIRE_KNOWN(wm::OBJ_OPTIMIZER, optimizer);
  const float v      = *static_cast<const float*>(value);
  const std::string k = parameter_name;
  auto& p            = optimizer->params;
  const auto t       = optimizer->type;
  /* which names each optimizer accepts: reference embedding_optimizer.cpp:119, :180-189, :303, :404-409 */
  bool adam = t == WHOLEMEMORY_OPT_LAZY_ADAM, ada = t == WHOLEMEMORY_OPT_ADAGRAD, rms = t == WHOLEMEMORY_OPT_RMSPROP;
  if (k == "weight_decay") p.weight_decay = v;
  else if (k == "epsilon" && (adam || ada || rms)) p.epsilon = v;
  else if (k == "beta1" && adam) p.beta1 = v;
  else if (k == "beta2" && adam) p.beta2 = v;
  else if (k == "adam_w" && adam) p.adam_w = v > 0.5f ? 1 : 0;
  else if (k == "alpha" && rms) p.alpha = v;
  else {
    WM_ERROR("parameter name %s is not valid for optimizer type %d", parameter_name, (int)t);
    return WHOLEMEMORY_INVALID_INPUT;
  }
  return WHOLEMEMORY_SUCCESS;
}

void wholememory_destroy_embedding_optimizer(wholememory_embedding_optimizer_t optimizer)
{
  if (!wm::obj_known(wm::OBJ_OPTIMIZER, optimizer)) return; /* null, or destroyed twice */
  wm::obj_unregister(wm::OBJ_OPTIMIZER, optimizer);
  delete optimizer;
}

wholememory_error_code_t wholememory_create_embedding_cache_policy(wholememory_embedding_cache_policy_t* cache_policy,
                                                                   wholememory_comm_t cache_level_comm,
                                                                   wholememory_memory_type_t memory_type,
                                                                   wholememory_memory_location_t memory_location,
                                                                   wholememory_access_type_t access_type,
                                                                   float cache_ratio)
{
  if (cache_policy == nullptr) return WHOLEMEMORY_INVALID_INPUT;
  if (cache_ratio > 1.0F || cache_ratio < 1.0F / 512) {
    WM_ERROR("cache_ratio should in range [1/512, 1.0]");
    return WHOLEMEMORY_INVALID_VALUE;
  }
  *cache_policy = new wholememory_embedding_cache_policy_{cache_level_comm, memory_type, memory_location, access_type, cache_ratio};
  wm::obj_register(wm::OBJ_CACHE_POLICY, *cache_policy);
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_destroy_embedding_cache_policy(wholememory_embedding_cache_policy_t cache_policy)
{
  if (cache_policy == nullptr) return WHOLEMEMORY_SUCCESS; /* deleting nothing succeeds, as with the reference's plain delete */
  WM_REQUIRE_KNOWN(wm::OBJ_CACHE_POLICY, cache_policy);
  wm::obj_unregister(wm::OBJ_CACHE_POLICY, cache_policy);
  delete cache_policy;
  return WHOLEMEMORY_SUCCESS;
}

wholememory_error_code_t wholememory_create_embedding(wholememory_embedding_t* wholememory_embedding,
                                                      wholememory_tensor_description_t* embedding_tensor_description,
                                                      wholememory_comm_t comm,
                                                      wholememory_memory_type_t memory_type,
                                                      wholememory_memory_location_t memory_location,
                                                      wholememory_embedding_cache_policy_t cache_policy,
                                                      size_t* embedding_entry_partition,
                                                      int user_defined_sms,
                                                      int round_robin_size)
{
  return wm::guarded("wholememory_create_embedding", [&]() -> wholememory_error_code_t {
    if (wholememory_embedding == nullptr || embedding_tensor_description == nullptr) return WHOLEMEMORY_INVALID_INPUT;
    WM_REQUIRE_LIVE(comm);
    if (cache_policy != nullptr) WM_REQUIRE_KNOWN(wm::OBJ_CACHE_POLICY, cache_policy);
    wholememory_matrix_description_t md;
    if (!wholememory_convert_tensor_desc_to_matrix(&md, embedding_tensor_description) || embedding_tensor_description->dim != 2) {
      WM_ERROR("wholememory_create_embedding input description must be 2D matrix");
      return WHOLEMEMORY_INVALID_INPUT;
    }
    if (cache_policy != nullptr) {
      WM_ERROR("cached embeddings are outside this build's scope: tables live in HBM on a B200 box (DESIGN.md)");
      return WHOLEMEMORY_NOT_IMPLEMENTED;
    }
    if (round_robin_size != 0 && embedding_entry_partition == nullptr) {
      WM_ERROR("round-robin sharding (round_robin_size=%d) is outside this build's scope", round_robin_size);
      return WHOLEMEMORY_NOT_IMPLEMENTED;
    }
    auto e        = std::make_unique<wholememory_embedding_>();
    e->comm       = comm;
    e->dtype      = md.dtype;
    /* reference set_gather_sms, embedding.cpp:442-455 */
    e->gather_sms = (user_defined_sms == -1 || (user_defined_sms > 0 && user_defined_sms <= 1568)) ? user_defined_sms : -1;
    wholememory_tensor_description_t padded;
    wholememory_copy_matrix_desc_to_tensor(&padded, &md);
    padded.storage_offset = 0;
    padded.strides[0]     = wm::pad_to_16_bytes(md.sizes[1], wholememory_dtype_get_element_size(md.dtype));
    WHOLEMEMORY_RETURN_ON_FAIL(wholememory_create_tensor(&e->allocated, &padded, comm, memory_type, memory_location, embedding_entry_partition));
    int64_t starts[2] = {0, 0}, ends[2] = {md.sizes[0], md.sizes[1]};
    wholememory_error_code_t rc = wholememory_tensor_get_subtensor(e->allocated, starts, ends, &e->user);
    if (rc != WHOLEMEMORY_SUCCESS) {
      (void)wholememory_destroy_tensor(e->allocated); /* collective like the creation: every rank takes this branch */
      return rc;
    }
    wm::obj_register(wm::OBJ_EMBEDDING, e.get());
    *wholememory_embedding = e.release();
    return WHOLEMEMORY_SUCCESS;
  });
}

wholememory_error_code_t wholememory_destroy_embedding(wholememory_embedding_t e)
{
  return wm::guarded("wholememory_destroy_embedding", [&]() -> wholememory_error_code_t {
    WM_REQUIRE_KNOWN(wm::OBJ_EMBEDDING, e);
    /* an embedding whose communicator is gone (wholememory_finalize) has lost its memory already: the tensor
     * objects below skip the collective free and only the host-side objects are released */
    wm::destroy_states(e);
    if (wm::live(e->comm)) wm::destroy_push_stage(&e->grad_stage);
    if (e->user) wholememory_destroy_tensor(e->user);
    if (e->allocated) WHOLEMEMORY_RETURN_ON_FAIL(wholememory_destroy_tensor(e->allocated));
    wm::obj_unregister(wm::OBJ_EMBEDDING, e);
    delete e;
    return WHOLEMEMORY_SUCCESS;
  });
}

wholememory_tensor_t wholememory_embedding_get_embedding_tensor(wholememory_embedding_t e)
{
  return wm::obj_known(wm::OBJ_EMBEDDING, e) ? e->user : nullptr;
}

wholememory_error_code_t wholememory_embedding_set_optimizer(wholememory_embedding_t e, wholememory_embedding_optimizer_t optimizer)
{
  return wm::guarded("wholememory_embedding_set_optimizer", [&]() -> wholememory_error_code_t {
    WM_REQUIRE_KNOWN(wm::OBJ_EMBEDDING, e);
    WM_REQUIRE_LIVE(e->comm);
    if (e->optimizer != nullptr) {
      WM_ERROR("optimizer can only be set once.");
      return WHOLEMEMORY_NOT_SUPPORTED;
    }
    if (optimizer == nullptr) return WHOLEMEMORY_SUCCESS;
    WM_REQUIRE_KNOWN(wm::OBJ_OPTIMIZER, optimizer);
    if (e->dtype != WHOLEMEMORY_DT_FLOAT) {
      WM_ERROR("Only float embedding supports training.");
      return WHOLEMEMORY_NOT_IMPLEMENTED;
    }
    e->optimizer = optimizer;
    return wm::create_states(e);
  });
}

wholememory_error_code_t wholememory_embedding_gather(wholememory_embedding_t e,
                                                      wholememory_tensor_t indices,
                                                      wholememory_tensor_t output,
                                                      bool /*adjust_cache*/,
                                                      wholememory_env_func_t* p_env_fns,
                                                      int64_t stream_int)
{
  WM_REQUIRE_KNOWN(wm::OBJ_EMBEDDING, e);
  /* the padded table is gathered, not the user view (reference noncached_embedding::gather, :553-562):
   * the output has D columns, so only the first D of each padded row are read. */
  return wholememory_gather(e->user, indices, output, p_env_fns, reinterpret_cast<void*>(stream_int), e->gather_sms);
}

wholememory_error_code_t wholememory_embedding_gather_gradient_apply(wholememory_embedding_t e,
                                                                     wholememory_tensor_t indices,
                                                                     wholememory_tensor_t grads,
                                                                     bool /*adjust_cache*/,
                                                                     float lr,
                                                                     wholememory_env_func_t* p_env_fns,
                                                                     int64_t stream_int)
{
  return wm::guarded("wholememory_embedding_gather_gradient_apply", [&]() -> wholememory_error_code_t {
    WM_REQUIRE_KNOWN(wm::OBJ_EMBEDDING, e);
    WM_REQUIRE_LIVE(e->allocated);
    WM_REQUIRE_LIVE(indices);
    WM_REQUIRE_LIVE(grads);
    if (e->optimizer != nullptr) WM_REQUIRE_KNOWN(wm::OBJ_OPTIMIZER, e->optimizer);
    return wm::gradient_apply(e, indices, grads, lr, p_env_fns, reinterpret_cast<cudaStream_t>(stream_int));
  });
}

const char* const* wholememory_embedding_get_optimizer_state_names(wholememory_embedding_t e)
{
  static const char* const none[] = {nullptr};
  if (!wm::obj_known(wm::OBJ_EMBEDDING, e) || !wm::obj_known(wm::OBJ_OPTIMIZER, e->optimizer)) return none;
  return e->optimizer->state_names.data();
}

wholememory_tensor_t wholememory_embedding_get_optimizer_state(wholememory_embedding_t e, const char* name)
{
  if (!wm::obj_known(wm::OBJ_EMBEDDING, e) || name == nullptr) return nullptr;
  for (auto& kv : e->named_states)
    if (kv.first == name) return kv.second;
  WM_ERROR("optimizer state name %s not found", name);
  return nullptr;
}

wholememory_error_code_t wholememory_embedding_writeback_cache(wholememory_embedding_t e, int64_t)
{
  return wm::obj_known(wm::OBJ_EMBEDDING, e) ? WHOLEMEMORY_SUCCESS : WHOLEMEMORY_INVALID_INPUT;
}

wholememory_error_code_t wholememory_embedding_drop_all_cache(wholememory_embedding_t e, int64_t)
{
  return wm::obj_known(wm::OBJ_EMBEDDING, e) ? WHOLEMEMORY_SUCCESS : WHOLEMEMORY_INVALID_INPUT;
}

} /* extern "C" */
