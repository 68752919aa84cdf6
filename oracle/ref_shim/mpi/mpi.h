/* mpi.h — single-process stand-in for the handful of MPI calls of the reference's mpi/ directory (Comm.cpp,
 * BoundaryExchange.cpp, MpiHelpers.hpp), so that the reference's own sources compile into oracle/_ref without an MPI
 * installation. One rank: every connection is local, no request is ever created with a non-zero count, and the
 * persistent-request calls are no-ops. TEST INFRASTRUCTURE. */
#ifndef REF_SHIM_MPI_H
#define REF_SHIM_MPI_H
#include <stdio.h>
#include <stdlib.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Request;
typedef int MPI_Datatype;
typedef int MPI_Fint;
typedef int MPI_Errhandler;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_DOUBLE 1
#define MPI_SUCCESS 0
#define MPI_MAX_ERROR_STRING 256
#define MPI_ERRORS_RETURN 1
#define MPI_ERRORS_ARE_FATAL 2
#define MPI_TAG_UB 32767
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_REQUEST_NULL 0
static inline int MPI_Initialized(int* flag) { *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_set_errhandler(MPI_Comm c, MPI_Errhandler e) { (void)c; (void)e; return MPI_SUCCESS; }
static inline MPI_Comm MPI_Comm_f2c(MPI_Fint f) { (void)f; return MPI_COMM_WORLD; }
static inline int MPI_Error_string(int code, char* s, int* len) { *len = snprintf(s, MPI_MAX_ERROR_STRING, "MPI stub error %d", code); return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return MPI_SUCCESS; }
static inline int MPI_Send_init(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* r) {
  (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; *r = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Recv_init(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* r) {
  (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; *r = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Startall(int n, MPI_Request* r) { (void)n; (void)r; return MPI_SUCCESS; }
static inline int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s) { (void)n; (void)r; (void)s; return MPI_SUCCESS; }
static inline int MPI_Request_free(MPI_Request* r) { *r = MPI_REQUEST_NULL; return MPI_SUCCESS; }
#ifdef __cplusplus
}
#endif
#endif
