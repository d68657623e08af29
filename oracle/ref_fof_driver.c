/* ref_fof_driver.c -- the reference's own primary friends-of-friends linking (fof_label_primary, fof.c:366-470) by
 * including fof.c where it lies (nothing is copied).  TEST INFRASTRUCTURE ONLY; pins oracle_fof_primary. */
#include <libgadget/fof.c>

void blackhole_make_one(int index, const double atime, const RandTable *const rnd) { endrun(1, "ref_fof_driver: blackhole_make_one reached\n"); }
int fof_save_particles(FOFGroups *fof, char *fname, int SaveParticles, Cosmology *CP, double atime, const double *MassTable, int MetalReturnOn,
                       MPI_Comm Comm) { endrun(1, "ref_fof_driver: fof_save_particles reached\n"); return 0; }

void ref_build_uniform_domain(DomainDecomp *d, int depth);
static DomainDecomp fof_dd;

/* MinID of every particle after the primary linking with linking length ll over particles of the types in mask */
int ref_fof_primary(int64_t n, const double *pos, const int64_t *ids, const unsigned char *type, double BoxSize, double ll, int64_t *minid_out)
{
    particle_alloc_memory(PartManager, BoxSize, n);
    PartManager->NumPart = n;
    DomainDecomp *dd = &fof_dd;
    ref_build_uniform_domain(dd, 0);
    for(int64_t i = 0; i < n; i++) {
        memset(&P[i], 0, sizeof(P[i]));
        for(int k = 0; k < 3; k++) P[i].Pos[k] = pos[3 * i + k];
        P[i].ID = ids[i]; P[i].Type = type[i]; P[i].Mass = 1; P[i].TopLeaf = 0;
    }
    set_fof_testpar(0, ll, 2);
    fof_init(1.0);
    ForceTree tree = {0};
    force_tree_rebuild_mask(&tree, dd, fof_params.FOFPrimaryLinkTypes, NULL);
    struct fof_particle_list *HaloLabel = (struct fof_particle_list *) mymalloc("HaloLabel", n * sizeof(struct fof_particle_list));
    fof_label_primary(HaloLabel, &tree, MPI_COMM_WORLD);
    for(int64_t i = 0; i < n; i++) minid_out[i] = HaloLabel[i].MinID;
    myfree(HaloLabel);
    force_tree_free(&tree);
    myfree(dd->Tasks); myfree(dd->TopLeaves); myfree(dd->TopNodes);
    memset(dd, 0, sizeof(*dd));
    myfree(P);
    return 0;
}
