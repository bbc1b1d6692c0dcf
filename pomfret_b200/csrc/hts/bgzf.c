/* hts-shim BGZF reader/writer on top of zlib (SAMv1 §4.1).
 * Single-threaded; bgzf_mt() is accepted and ignored. */
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <zlib.h>
#include "htslib/bgzf.h"

#define BLOCK_HEADER_LENGTH 18
#define BLOCK_FOOTER_LENGTH 8

static const uint8_t k_eof_block[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C',
                                       2,    0,    27, 0, 3, 0, 0, 0, 0, 0, 0,    0, 0, 0};

static inline uint16_t ld16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t ld32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline void st16(uint8_t *p, uint16_t v) { p[0] = v & 0xff; p[1] = v >> 8; }
static inline void st32(uint8_t *p, uint32_t v) {
    p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; p[2] = (v >> 16) & 0xff; p[3] = (v >> 24) & 0xff;
}

BGZF *bgzf_open(const char *path, const char *mode) {
    int is_write = strchr(mode, 'w') != NULL;
    FILE *f = fopen(path, is_write ? "wb" : "rb");
    if (!f) return NULL;
    BGZF *fp = (BGZF *)calloc(1, sizeof(BGZF));
    fp->fp = f;
    fp->is_write = is_write;
    fp->compress_level = -1;
    for (const char *p = mode; *p; p++) {
        if (*p >= '0' && *p <= '9') fp->compress_level = *p - '0';
        if (*p == 'u') fp->compress_level = 0;
    }
    fp->ublock = (uint8_t *)malloc(BGZF_MAX_BLOCK_SIZE);
    fp->cblock = (uint8_t *)malloc(BGZF_MAX_BLOCK_SIZE);
    setvbuf(f, NULL, _IOFBF, 1 << 20);
    return fp;
}

int bgzf_mt(BGZF *fp, int n_threads, int n_sub_blks) {
    (void)fp; (void)n_threads; (void)n_sub_blks;
    return 0;
}

/* ---------- reading ---------- */

/* Load the block at the current file position. Returns 0 ok (block_length==0 at EOF), -1 error. */
static int read_block(BGZF *fp) {
    uint8_t hdr[BLOCK_HEADER_LENGTH];
    int64_t addr = ftello(fp->fp);
    size_t got = fread(hdr, 1, BLOCK_HEADER_LENGTH, fp->fp);
    if (got == 0) {
        fp->block_length = 0;
        fp->block_offset = 0;
        fp->block_address = addr;
        fp->at_eof = 1;
        return 0;
    }
    if (got != BLOCK_HEADER_LENGTH || hdr[0] != 0x1f || hdr[1] != 0x8b || hdr[2] != 8 || !(hdr[3] & 4)) {
        fp->errcode |= 1;
        return -1;
    }
    int xlen = ld16(hdr + 10);
    /* the first 6 bytes of the extra field were read with the header */
    int bsize = -1;
    uint8_t *extra = fp->cblock;
    memcpy(extra, hdr + 12, 6);
    if (xlen > 6) {
        if (fread(extra + 6, 1, xlen - 6, fp->fp) != (size_t)(xlen - 6)) { fp->errcode |= 1; return -1; }
    } else if (xlen < 6) {
        fp->errcode |= 1;
        return -1;
    }
    for (int off = 0; off + 4 <= xlen;) {
        int slen = ld16(extra + off + 2);
        if (extra[off] == 'B' && extra[off + 1] == 'C' && slen == 2) bsize = ld16(extra + off + 4);
        off += 4 + slen;
    }
    if (bsize < 0) { fp->errcode |= 1; return -1; }
    int remaining = bsize + 1 - 12 - xlen; /* cdata + footer */
    if (remaining < BLOCK_FOOTER_LENGTH) { fp->errcode |= 1; return -1; }
    if (fread(fp->cblock, 1, remaining, fp->fp) != (size_t)remaining) { fp->errcode |= 1; return -1; }
    int clen = remaining - BLOCK_FOOTER_LENGTH;
    uint32_t isize = ld32(fp->cblock + clen + 4);
    if (isize > BGZF_MAX_BLOCK_SIZE) { fp->errcode |= 1; return -1; }
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    zs.next_in = fp->cblock;
    zs.avail_in = clen;
    zs.next_out = fp->ublock;
    zs.avail_out = BGZF_MAX_BLOCK_SIZE;
    if (inflateInit2(&zs, -15) != Z_OK) { fp->errcode |= 2; return -1; }
    int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || zs.total_out != isize) { fp->errcode |= 2; return -1; }
    /* the footer's CRC32 covers the inflated block (htslib checks it too) */
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), fp->ublock, isize) != ld32(fp->cblock + clen)) { fp->errcode |= 2; return -1; }
    fp->block_address = addr;
    fp->next_address = addr + bsize + 1;
    fp->block_length = (int)isize;
    fp->block_offset = 0;
    return 0;
}

ssize_t bgzf_read(BGZF *fp, void *data, size_t length) {
    uint8_t *out = (uint8_t *)data;
    size_t done = 0;
    while (done < length) {
        int avail = fp->block_length - fp->block_offset;
        if (avail <= 0) {
            if (read_block(fp) != 0) return -1;
            avail = fp->block_length - fp->block_offset;
            if (avail <= 0) {
                if (fp->at_eof) break;
                continue; /* empty block inside the file */
            }
        }
        size_t n = length - done < (size_t)avail ? length - done : (size_t)avail;
        memcpy(out + done, fp->ublock + fp->block_offset, n);
        fp->block_offset += (int)n;
        done += n;
    }
    if (fp->block_offset == fp->block_length && fp->block_length > 0) {
        /* normalise so that tell() names the start of the next block */
        fp->block_address = fp->next_address;
        fp->block_offset = fp->block_length = 0;
    }
    return (ssize_t)done;
}

int64_t bgzf_tell(BGZF *fp) {
    if (fp->is_write) return ((int64_t)ftello(fp->fp) << 16) | (fp->block_offset & 0xffff);
    return (fp->block_address << 16) | (fp->block_offset & 0xffff);
}

int64_t bgzf_seek(BGZF *fp, int64_t pos, int whence) {
    if (fp->is_write || whence != SEEK_SET) { fp->errcode |= 4; return -1; }
    int64_t addr = pos >> 16;
    int off = (int)(pos & 0xffff);
    if (fseeko(fp->fp, addr, SEEK_SET) != 0) { fp->errcode |= 4; return -1; }
    fp->at_eof = 0;
    fp->block_length = 0;
    fp->block_offset = 0;
    fp->block_address = addr;
    if (off > 0 || 1) {
        if (read_block(fp) != 0) return -1;
        if (off > fp->block_length) { fp->errcode |= 4; return -1; }
        fp->block_offset = off;
        if (fp->block_length > 0 && fp->block_offset == fp->block_length) {
            fp->block_address = fp->next_address;
            fp->block_offset = fp->block_length = 0;
        }
    }
    return 0;
}

/* ---------- writing ---------- */

/* one BGZF block around `ulen` bytes of `src`: header with BSIZE, raw deflate, CRC-32, ISIZE; dst holds
 * BGZF_MAX_BLOCK_SIZE bytes.  Returns the block's size or -1.  (Also what a caller that compresses blocks on its own
 * threads uses, so that its blocks are the bytes bgzf_write would have produced.) */
int pomfret_bgzf_compress_block(uint8_t *dst, const uint8_t *src, int ulen, int compress_level) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    zs.next_in = (Bytef *)src;
    zs.avail_in = ulen;
    zs.next_out = dst + BLOCK_HEADER_LENGTH;
    zs.avail_out = BGZF_MAX_BLOCK_SIZE - BLOCK_HEADER_LENGTH - BLOCK_FOOTER_LENGTH;
    int level = compress_level < 0 ? Z_DEFAULT_COMPRESSION : compress_level;
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    int rc = deflate(&zs, Z_FINISH);
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return -1;
    int clen = (int)zs.total_out;
    int total = clen + BLOCK_HEADER_LENGTH + BLOCK_FOOTER_LENGTH;
    memcpy(dst, k_eof_block, BLOCK_HEADER_LENGTH);
    st16(dst + 16, (uint16_t)(total - 1));
    uint32_t crc = (uint32_t)crc32(crc32(0L, NULL, 0), src, ulen);
    st32(dst + BLOCK_HEADER_LENGTH + clen, crc);
    st32(dst + BLOCK_HEADER_LENGTH + clen + 4, (uint32_t)ulen);
    return total;
}

int pomfret_bgzf_eof_block(const uint8_t **p) {
    *p = k_eof_block;
    return (int)sizeof(k_eof_block);
}

static int deflate_block(BGZF *fp, int ulen) {
    return pomfret_bgzf_compress_block(fp->cblock, fp->ublock, ulen, fp->compress_level);
}

int bgzf_flush(BGZF *fp) {
    if (!fp->is_write) return 0;
    while (fp->block_offset > 0) {
        int n = fp->block_offset;
        int total = deflate_block(fp, n);
        if (total < 0) { fp->errcode |= 2; return -1; }
        if (fwrite(fp->cblock, 1, total, fp->fp) != (size_t)total) { fp->errcode |= 8; return -1; }
        fp->block_offset = 0;
    }
    return 0;
}

int bgzf_flush_try(BGZF *fp, ssize_t size) {
    if (fp->block_offset + size > BGZF_BLOCK_SIZE) return bgzf_flush(fp);
    return 0;
}

ssize_t bgzf_write(BGZF *fp, const void *data, size_t length) {
    const uint8_t *in = (const uint8_t *)data;
    size_t done = 0;
    while (done < length) {
        size_t room = BGZF_BLOCK_SIZE - fp->block_offset;
        size_t n = length - done < room ? length - done : room;
        memcpy(fp->ublock + fp->block_offset, in + done, n);
        fp->block_offset += (int)n;
        done += n;
        if (fp->block_offset == BGZF_BLOCK_SIZE && bgzf_flush(fp) != 0) return -1;
    }
    return (ssize_t)done;
}

int bgzf_close(BGZF *fp) {
    if (!fp) return -1;
    int rc = 0;
    if (fp->is_write) {
        if (bgzf_flush(fp) != 0) rc = -1;
        if (fwrite(k_eof_block, 1, sizeof(k_eof_block), fp->fp) != sizeof(k_eof_block)) rc = -1;
    }
    if (fclose(fp->fp) != 0) rc = -1;
    free(fp->ublock);
    free(fp->cblock);
    free(fp);
    return rc;
}
