// C entry points of the host front end used by the Python tests and bench.py (ctypes).
#include <cstring>
#include "loader.h"

using namespace pomfret;

extern "C" {

void *pomfret_host_bam_open(const char *path) {
    BamReader *r = new BamReader();
    if (!r->open(path)) { delete r; return nullptr; }
    return r;
}
void pomfret_host_bam_close(void *h) { delete (BamReader *)h; }

void *pomfret_host_window_load(void *bam, const char *chrom, uint32_t ref_start, uint32_t ref_end,
                               int readlen_threshold, int min_mapq, int *rc_out) {
    WindowReads *w = new WindowReads();
    int rc = load_window(*(BamReader *)bam, chrom, ref_start, ref_end, readlen_threshold, min_mapq, nullptr, w);
    if (rc_out) *rc_out = rc;
    if (rc != 0) { delete w; return nullptr; }
    return w;
}
int pomfret_host_window_n(void *w) { return (int)((WindowReads *)w)->descs.size(); }
const pomfret_gpu_read_desc *pomfret_host_window_descs(void *w) { return ((WindowReads *)w)->descs.data(); }
const char *pomfret_host_window_qname(void *w, int i) { return ((WindowReads *)w)->qname((size_t)i); }
uint64_t pomfret_host_window_bases(void *w) { return ((WindowReads *)w)->n_bases; }
void pomfret_host_window_free(void *w) { delete (WindowReads *)w; }

}  // extern "C"
