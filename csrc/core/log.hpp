// Logging and assertion macros.
// Behavioural parity with reference src/log.hpp:35-98 (levels ERROR=0, INFO, DEBUG, TRACE chosen by
// MLSL_LOG_LEVEL; a failed assertion prints, tears the environment down and terminates).  Ours adds a
// "throw" mode (MLSL_ASSERT_MODE=throw, or set programmatically) so that error paths are unit-testable and
// language bindings can surface failures as exceptions instead of killing the interpreter.
#pragma once
#include <cstdarg>
#include <stdexcept>
#include <string>

namespace mlslb {

enum LogLevel { LOG_ERROR = 0, LOG_INFO = 1, LOG_DEBUG = 2, LOG_TRACE = 3 };

struct Error : public std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

int log_level();
void set_log_level(int lvl);
void set_assert_throws(bool on);
bool assert_throws();
void log_emit(int lvl, const char* file, int line, const char* func, const char* fmt, ...)
    __attribute__((format(printf, 5, 6)));
[[noreturn]] void fail(const char* file, int line, const char* func, const char* cond, const char* fmt, ...)
    __attribute__((format(printf, 5, 6)));
void set_fail_hook(void (*hook)());   // called once before process termination (fail-fast teardown)
void set_log_rank(int rank);

}  // namespace mlslb

#define MLSLB_LOG(lvl, ...)                                                            \
  do {                                                                                 \
    if ((int)(lvl) <= ::mlslb::log_level())                                            \
      ::mlslb::log_emit((int)(lvl), __FILE__, __LINE__, __func__, __VA_ARGS__);        \
  } while (0)

#define MLSLB_ASSERT(cond, ...)                                                        \
  do {                                                                                 \
    if (__builtin_expect(!(cond), 0))                                                  \
      ::mlslb::fail(__FILE__, __LINE__, __func__, #cond, __VA_ARGS__);                 \
  } while (0)

#define MLSLB_CUDA(expr)                                                               \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess)                                                             \
      ::mlslb::fail(__FILE__, __LINE__, __func__, #expr, "CUDA error %d: %s", (int)_e, \
                    cudaGetErrorString(_e));                                           \
  } while (0)
