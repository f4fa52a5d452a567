#include "common.cuh"
#include <stdarg.h>

namespace spk {
static thread_local char t_err[1024] = {0};
char* last_error_buf() { return t_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}
}  // namespace spk
