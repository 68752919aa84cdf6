/* gptl.h — the two GPTL timer calls the reference's profiling.hpp wraps, as no-ops. TEST INFRASTRUCTURE. */
#ifndef REF_SHIM_GPTL_H
#define REF_SHIM_GPTL_H
#ifdef __cplusplus
extern "C" {
#endif
static inline int GPTLstart(const char* name) { (void)name; return 0; }
static inline int GPTLstop(const char* name) { (void)name; return 0; }
#ifdef __cplusplus
}
#endif
#endif
