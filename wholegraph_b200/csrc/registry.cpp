/*
 * Registry of the opaque objects handed out across the C ABI (communicators, memory handles, tensors, embeddings,
 * optimizers, cache policies).  Every extern "C" entry point looks its handle arguments up here before touching
 * them, so a stale handle -- e.g. a communicator that wholememory_finalize() already destroyed
 * (reference cpp/src/wholememory/initialize.cpp:73-77 destroys every communicator there, and a dying communicator
 * takes its memory handles with it, communicator.cpp:808-827) -- comes back as WHOLEMEMORY_INVALID_INPUT instead of
 * undefined behaviour.  The reference does not check; this is strictly stronger and costs one hash lookup per call.
 */
#include "wm_internal.hpp"

#include <shared_mutex>
#include <unordered_set>

namespace wm {

namespace {
struct object_table {
  std::shared_mutex mu;
  std::unordered_set<const void*> live[OBJ_KINDS];
};
object_table& table()
{
  static object_table* t = new object_table(); /* never destroyed: entry points may run during process teardown */
  return *t;
}
}  // namespace

void obj_register(obj_kind k, const void* p)
{
  auto& t = table();
  std::unique_lock<std::shared_mutex> lk(t.mu);
  t.live[k].insert(p);
}

void obj_unregister(obj_kind k, const void* p)
{
  auto& t = table();
  std::unique_lock<std::shared_mutex> lk(t.mu);
  t.live[k].erase(p);
}

bool obj_known(obj_kind k, const void* p)
{
  if (p == nullptr) return false;
  auto& t = table();
  std::shared_lock<std::shared_mutex> lk(t.mu);
  return t.live[k].count(p) != 0;
}

std::string api_name(const char* pretty)
{
  std::string s(pretty);
  size_t paren = s.find('(');
  if (paren == std::string::npos) return s;
  size_t start = s.rfind(' ', paren);
  return s.substr(start == std::string::npos ? 0 : start + 1, paren - (start == std::string::npos ? 0 : start + 1));
}

bool live(wholememory_comm_t c) { return obj_known(OBJ_COMM, c); }

bool live(wholememory_handle_t h) { return obj_known(OBJ_HANDLE, h) && obj_known(OBJ_COMM, h->comm); }

bool live(wholememory_tensor_t t)
{
  if (!obj_known(OBJ_TENSOR, t)) return false;
  if (t->root != t && !obj_known(OBJ_TENSOR, t->root)) return false; /* a view outlived the tensor it was cut from */
  return !t->is_wm || live(t->handle);
}

}  // namespace wm
