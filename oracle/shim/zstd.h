// Stand-in for <zstd.h> (the development header is absent from this image; libzstd.so.1 is present): the handful of
// declarations of zstd's stable public API that the reference's Common/BinaryIO.cpp uses, written from the zstd manual.
// TEST INFRASTRUCTURE ONLY — lets oracle/_ref compile the reference's own (de)serialiser and link the system libzstd.
#pragma once
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct ZSTD_CCtx_s ZSTD_CCtx;
typedef struct ZSTD_DCtx_s ZSTD_DCtx;
typedef struct ZSTD_inBuffer_s {
    const void* src;
    size_t size;
    size_t pos;
} ZSTD_inBuffer;
typedef struct ZSTD_outBuffer_s {
    void* dst;
    size_t size;
    size_t pos;
} ZSTD_outBuffer;
typedef enum { ZSTD_c_compressionLevel = 100, ZSTD_c_checksumFlag = 201 } ZSTD_cParameter;
typedef enum { ZSTD_e_continue = 0, ZSTD_e_flush = 1, ZSTD_e_end = 2 } ZSTD_EndDirective;
#define ZSTD_CLEVEL_DEFAULT 3
ZSTD_CCtx* ZSTD_createCCtx(void);
size_t ZSTD_freeCCtx(ZSTD_CCtx* cctx);
size_t ZSTD_CCtx_setParameter(ZSTD_CCtx* cctx, ZSTD_cParameter param, int value);
size_t ZSTD_compressStream2(ZSTD_CCtx* cctx, ZSTD_outBuffer* output, ZSTD_inBuffer* input, ZSTD_EndDirective endOp);
ZSTD_DCtx* ZSTD_createDCtx(void);
size_t ZSTD_freeDCtx(ZSTD_DCtx* dctx);
size_t ZSTD_decompressStream(ZSTD_DCtx* zds, ZSTD_outBuffer* output, ZSTD_inBuffer* input);
unsigned ZSTD_isError(size_t code);
#ifdef __cplusplus
}
#endif
