/* oracle_fof.c -- CPU restatement of the primary friends-of-friends linking of the reference
 * (fof_label_primary + fof_primary_ngbiter, libgadget/fof.c:366-470,540-579): particles of the primary link types
 * closer than the linking length (periodic, NEAREST) belong to one group, and every member carries the smallest
 * particle ID of its group (HaloLabel[].MinID); particles of other types keep their own ID.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PINNED against the reference's own fof.c compiled unmodified
 * (oracle/_ref/libref_domain.so, tests/golden/ref_fof.npz).
 *
 * The reference reaches the fixed point by repeated tree walks that merge heads and propagate MinID; the result is the
 * connected components of the distance graph, computed here with a cell list and union-find. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static int64_t uf_find(int64_t *parent, int64_t i)
{
    while(parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; }
    return i;
}
static void uf_union(int64_t *parent, int64_t a, int64_t b)
{
    a = uf_find(parent, a); b = uf_find(parent, b);
    if(a == b) return;
    if(a < b) parent[b] = a; else parent[a] = b;
}

/* mask: bit t set = particle type t is a primary link type (FOFPrimaryLinkTypes).  Returns 0, 1 out of memory. */
int oracle_fof_primary(int64_t n, const double *pos, const int64_t *ids, const uint8_t *type, int mask, double BoxSize, double ll,
                       int64_t *minid)
{
    int nc = (int) floor(BoxSize / ll);
    if(nc > 256) nc = 256;
    if(nc < 1) nc = 1;
    const double cs = BoxSize / nc;
    const int64_t ncell = (int64_t) nc * nc * nc;
    int64_t *head = (int64_t *) malloc(sizeof(int64_t) * ncell), *next = (int64_t *) malloc(sizeof(int64_t) * (n > 0 ? n : 1));
    int64_t *parent = (int64_t *) malloc(sizeof(int64_t) * (n > 0 ? n : 1));
    int32_t *cell = (int32_t *) malloc(sizeof(int32_t) * 3 * (n > 0 ? n : 1));
    if(!head || !next || !parent || !cell) { free(head); free(next); free(parent); free(cell); return 1; }
    for(int64_t c = 0; c < ncell; c++) head[c] = -1;
    for(int64_t i = 0; i < n; i++) {
        parent[i] = i; next[i] = -1;
        if(!((mask >> type[i]) & 1)) continue;
        for(int k = 0; k < 3; k++) {
            int c = (int) floor(pos[3 * i + k] / cs);
            c %= nc; if(c < 0) c += nc;
            cell[3 * i + k] = c;
        }
        const int64_t ci = ((int64_t) cell[3 * i] * nc + cell[3 * i + 1]) * nc + cell[3 * i + 2];
        next[i] = head[ci]; head[ci] = i;
    }
    const int reach = nc >= 3 ? 1 : 0;          /* with fewer than 3 cells per side every pair is examined once via cell (0,0,0) */
    for(int64_t i = 0; i < n; i++) {
        if(!((mask >> type[i]) & 1)) continue;
        for(int dx = -reach; dx <= reach; dx++) for(int dy = -reach; dy <= reach; dy++) for(int dz = -reach; dz <= reach; dz++) {
            int cx, cy, cz;
            if(reach) {
                cx = (cell[3 * i] + dx + nc) % nc; cy = (cell[3 * i + 1] + dy + nc) % nc; cz = (cell[3 * i + 2] + dz + nc) % nc;
            } else { cx = cy = cz = -1; }
            for(int64_t c = reach ? ((int64_t) cx * nc + cy) * nc + cz : 0; c < (reach ? ((int64_t) cx * nc + cy) * nc + cz + 1 : ncell); c++)
                for(int64_t j = head[c]; j >= 0; j = next[j]) {
                    if(j <= i) continue;
                    double r2 = 0;
                    for(int k = 0; k < 3; k++) {
                        double d = pos[3 * i + k] - pos[3 * j + k];
                        if(d > 0.5 * BoxSize) d -= BoxSize;              /* NEAREST */
                        if(d < -0.5 * BoxSize) d += BoxSize;
                        r2 += d * d;
                    }
                    if(r2 <= ll * ll) uf_union(parent, i, j);            /* treewalk.c:989-993: r2 <= h2 for an asymmetric search */
                }
        }
    }
    for(int64_t i = 0; i < n; i++) minid[i] = ids[i];
    for(int64_t i = 0; i < n; i++) { const int64_t r = uf_find(parent, i); if(ids[i] < minid[r]) minid[r] = ids[i]; }
    for(int64_t i = 0; i < n; i++) minid[i] = minid[uf_find(parent, i)];
    free(head); free(next); free(parent); free(cell);
    return 0;
}
