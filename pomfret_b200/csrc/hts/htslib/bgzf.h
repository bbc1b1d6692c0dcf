/* hts-shim BGZF layer (SAMv1 §4.1: concatenated gzip members with a BC extra
 * field holding the block size; virtual offset = coffset<<16 | uoffset). */
#ifndef POMFRET_HTS_SHIM_BGZF_H
#define POMFRET_HTS_SHIM_BGZF_H
#include <stdint.h>
#include <stdio.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGZF_MAX_BLOCK_SIZE 0x10000
#define BGZF_BLOCK_SIZE 0xff00

typedef struct BGZF {
    FILE *fp;
    int is_write;
    int compress_level;
    int errcode;
    int at_eof;
    /* current uncompressed block */
    uint8_t *ublock;
    int block_length;      /* valid bytes in ublock (read) / filled bytes (write) */
    int block_offset;      /* read cursor in ublock */
    int64_t block_address; /* file offset of the current block */
    int64_t next_address;  /* file offset of the block after the current one (read) */
    uint8_t *cblock;
} BGZF;

BGZF *bgzf_open(const char *path, const char *mode);
int bgzf_close(BGZF *fp);
ssize_t bgzf_read(BGZF *fp, void *data, size_t length);
ssize_t bgzf_write(BGZF *fp, const void *data, size_t length);
int64_t bgzf_tell(BGZF *fp);
int64_t bgzf_seek(BGZF *fp, int64_t pos, int whence);
int bgzf_flush(BGZF *fp);
int bgzf_flush_try(BGZF *fp, ssize_t size);
int bgzf_mt(BGZF *fp, int n_threads, int n_sub_blks);
/* shim extras: a whole BGZF block from ulen bytes (what bgzf_write emits), the EOF marker block */
int pomfret_bgzf_compress_block(uint8_t *dst, const uint8_t *src, int ulen, int compress_level);
int pomfret_bgzf_eof_block(const uint8_t **p);

#ifdef __cplusplus
}
#endif
#endif
