/* EPLIB-style entry points on the mlsl-b200 runtime.
 *
 * The reference ships its endpoint-proxy runtime as a library of its own (reference eplib/eplib.h): explicit init /
 * teardown, allocation in the memory its servers can reach, suspending and resuming the servers, and file reads executed by
 * a server while the caller computes.  Programs written against that header find the same calls here, implemented by the
 * progress threads and the symmetric heap of this library.  What is missing on purpose: the MPI types - there is no MPI
 * underneath, so requests are EPLIB_Request handles instead of MPI_Request, streams are EPLIB_FILE instead of FILE* (the
 * reference's FILE* belongs to the server process and is just as opaque to the caller), and MPI_Comm_create_endpoints /
 * EPLIB_split_comm / EPLIB_comm_set_info have no counterpart (groups are Distributions; include/mlsl.h).
 * `epid` (which endpoint serves the call) is accepted and ignored: the progress engine places the work.
 */
#ifndef MLSL_B200_EPLIB_H_
#define MLSL_B200_EPLIB_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eplib_file_s* EPLIB_FILE;
typedef unsigned long long EPLIB_Request;

/* Init / teardown (reference eplib/eplib.h:36-38).  EPLIB_init initialises the library unless the program already did
 * (Environment::Init); EPLIB_finalize undoes only what EPLIB_init did.  Both return 0 on success. */
int EPLIB_init(void);
int EPLIB_finalize(void);

/* Memory the servers - here: the peers' kernels / the other ranks - can address (reference eplib/eplib.h:40-48) */
void* EPLIB_malloc(size_t bytes);
void* EPLIB_realloc(void* ptr, size_t bytes);
void* EPLIB_calloc(size_t count, size_t size);
void* EPLIB_memalign(size_t alignment, size_t bytes);
void EPLIB_free(void* ptr);
int EPLIB_memory_is_shmem(void* ptr);          /* 1: a collective uses it in place, 0: it would be staged */
void EPLIB_set_mem_hooks(void);                /* accepted; malloc itself is never redirected */
void* EPLIB_quant_params_submit(void* mlsl_quant_params);   /* a mlsl_quant_params* (include/mlsl.h); returns its argument */

/* Server management (reference eplib/eplib.h:50-52): EPLIB_suspend parks the progress threads until EPLIB_execute */
void EPLIB_execute(void);
void EPLIB_suspend(void);

/* File reads on a progress thread (reference eplib/eplib.h:54-61).  Streams are read front to back like a FILE*. */
EPLIB_FILE EPLIB_fopen(int epid, const char* filename, const char* mode);                 /* mode must start with 'r' */
size_t EPLIB_fread(int epid, void* buffer, size_t size, size_t count, EPLIB_FILE stream);   /* blocking; items read */
size_t EPLIB_fread_nb(int epid, void* buffer, size_t size, size_t count, EPLIB_FILE stream, EPLIB_Request* request);
/* open + read + close in one non-blocking command */
size_t EPLIB_forc_nb(int epid, const char* filename, const char* mode, void* buffer, size_t size, size_t count, EPLIB_Request* request);
int EPLIB_fwait(EPLIB_Request* request, size_t* readcount);                 /* readcount: ITEMS read, as with fread */
int EPLIB_fwaitall(int count, EPLIB_Request* requests, size_t* readcounts);
int EPLIB_fclose(int epid, EPLIB_FILE stream);

#ifdef __cplusplus
}
#endif
#endif
