// Error text (thread-local) and the optional launch profiler of libspkemb.so.
#include "common.cuh"
#include <stdarg.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace spk {
static thread_local char t_err[1024] = {0};
char* last_error_buf() { return t_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// Profiler: when enabled, every launcher brackets its kernel with CUDA events on the launching
// stream and records its algorithmic FLOPs / bytes.  bench.py uses it for the per-kernel roofline.
struct ProfRec {
  const char* tag;
  cudaEvent_t e0, e1;
  double flops, bytes;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

bool prof_enabled() { return g_prof_on; }

void prof_set(bool on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on;
}

int prof_begin(const char* tag, double flops, double bytes, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.tag = tag; r.flops = flops; r.bytes = bytes;
  r.e0 = prof_event();
  r.e1 = prof_event();
  cudaEventRecord(r.e0, st);
  g_prof_recs.push_back(r);
  return static_cast<int>(g_prof_recs.size()) - 1;
}

void prof_end(int idx, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx >= 0 && idx < static_cast<int>(g_prof_recs.size())) cudaEventRecord(g_prof_recs[idx].e1, st);
}

// Aggregates and clears the records: one line per tag "tag count total_ms flops bytes".
int prof_report(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Agg { long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  for (ProfRec& r : g_prof_recs) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (!agg.count(r.tag)) order.push_back(r.tag);
    Agg& a = agg[r.tag];
    a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
    g_prof_pool.push_back(r.e0);
    g_prof_pool.push_back(r.e1);
  }
  g_prof_recs.clear();
  size_t off = 0;
  if (cap) buf[0] = 0;
  for (const std::string& t : order) {
    const Agg& a = agg[t];
    int n = snprintf(buf + off, off < cap ? cap - off : 0, "%s %ld %.6f %.6e %.6e\n", t.c_str(), a.n, a.ms, a.flops,
                     a.bytes);
    if (n < 0 || off + n >= cap) break;
    off += n;
  }
  return static_cast<int>(off);
}
}  // namespace spk
