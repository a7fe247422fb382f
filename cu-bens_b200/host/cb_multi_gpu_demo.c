/* cb_multi_gpu_demo.c - a C host running the element-partitioned hot path on several GPUs with no Python,
 * no torch and no MPI: one process per GPU, the ncclUniqueId travels through a file, the per-iteration
 * collective is the library's own (cb_comm_init / cb_residual_allreduce, include/cubens_b200.h).
 *
 *   cb_multi_gpu_demo <model.bin> <rank> <world> <uid-file> <out.bin>
 *
 * model.bin (written by cubens_b200.partition.write_submodel): the sub-model of this rank in the host layout
 * of cb_model - global joint / equation numbering, its own elements plus the halo elements - followed by the
 * owned joint range, the load vector q, one displacement increment dd and the load factor.  The program does
 * what one Newton iteration of main.c:1891-2030 does on the device path - cb_begin_increment, cb_stiff,
 * cb_update_forces, cb_residual_sums + cb_residual_allreduce (NCCL over NVLink) - and writes the eleven
 * all-reduced sums (and 1 / 0: exchanged over mapped peer memory / NCCL), so that a test can hold them against a one-GPU run of the whole model.                 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "../../include/cubens_b200.h"

static void *rd(FILE *f, long *count, size_t elem)
{
    long n = 0;
    if (fread(&n, sizeof n, 1, f) != 1) { fprintf(stderr, "short read\n"); exit(2); }
    *count = n;
    if (n == 0) return NULL;
    void *p = malloc((size_t)n * elem);
    if (!p || fread(p, elem, (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

int main(int argc, char **argv)
{
    if (argc != 6) { fprintf(stderr, "usage: %s model.bin rank world uid-file out.bin\n", argv[0]); return 2; }
    const int rank = atoi(argv[2]), world = atoi(argv[3]);
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    cb_sizes sz; cb_flags fl; long own[2], n;
    if (fread(&sz, sizeof sz, 1, f) != 1 || fread(&fl, sizeof fl, 1, f) != 1 || fread(own, sizeof own, 1, f) != 1) return 2;
    fl.device = rank % (cb_device_count() > 0 ? cb_device_count() : 1);
    cb_model m; memset(&m, 0, sizeof m);
    m.x = rd(f, &n, 8); m.minc = rd(f, &n, 8); m.jcode = rd(f, &n, 8); m.mcode = rd(f, &n, 8); m.maxa = rd(f, &n, 8);
    m.emod = rd(f, &n, 8); m.dens = rd(f, &n, 8); m.carea = rd(f, &n, 8); m.llength = rd(f, &n, 8);
    m.c1 = rd(f, &n, 8); m.c2 = rd(f, &n, 8); m.c3 = rd(f, &n, 8);
    m.nu = rd(f, &n, 8); m.thick = rd(f, &n, 8); m.farea = rd(f, &n, 8); m.slength = rd(f, &n, 8); m.xlocal = rd(f, &n, 8);
    m.gmod = rd(f, &n, 8); m.istrong = rd(f, &n, 8); m.iweak = rd(f, &n, 8); m.ipolar = rd(f, &n, 8); m.iwarp = rd(f, &n, 8);
    m.auxpt = rd(f, &n, 8); m.offset = rd(f, &n, 8); m.osflag = rd(f, &n, 4); m.mendrel = rd(f, &n, 4); m.efFE_ref = rd(f, &n, 8);
    m.yield = rd(f, &n, 8); m.zstrong = rd(f, &n, 8); m.zweak = rd(f, &n, 8);
    m.nnorm = rd(f, &n, 8); m.tarea = rd(f, &n, 8); m.fdens = rd(f, &n, 8);
    double *q = rd(f, &n, 8), *dd = rd(f, &n, 8), lpf = 0;
    if (fread(&lpf, sizeof lpf, 1, f) != 1) return 2;
    fclose(f);

    cb_handle *h = NULL;
#define CK(call) do { int rc_ = (call); if (rc_) { fprintf(stderr, "rank %d: %s -> %d: %s\n", rank, #call, rc_, cb_last_error()); return 1; } } while (0)
    CK(cb_create(&sz, &fl, &m, &h));
    CK(cb_set_owned_joints(h, own[0], own[1]));
    /* the 128-byte ncclUniqueId: rank 0 writes it (atomically, via rename), the others wait for the file */
    unsigned char uid[128];
    if (rank == 0) {
        CK(cb_comm_unique_id(uid));
        char tmp[4096]; snprintf(tmp, sizeof tmp, "%s.tmp", argv[4]);
        FILE *u = fopen(tmp, "wb"); if (!u) { perror(tmp); return 2; }
        fwrite(uid, 1, sizeof uid, u); fclose(u); rename(tmp, argv[4]);
    } else {
        FILE *u = NULL;
        for (int tries = 0; tries < 600 && !(u = fopen(argv[4], "rb")); ++tries) usleep(100000);
        if (!u || fread(uid, 1, sizeof uid, u) != sizeof uid) { fprintf(stderr, "rank %d: no unique id\n", rank); return 2; }
        fclose(u);
    }
    CK(cb_comm_init(h, uid, rank, world));
    CK(cb_begin_increment(h));
    CK(cb_stiff(h, CB_GEN_IP));
    double dlpf = 1.0; int fr = 0, sh = 0;
    double *ftmp = calloc((size_t)sz.NEQ, sizeof(double));
    CK(cb_update_forces(h, dd, &dlpf, 0, ftmp, &fr, &sh));
    CK(cb_set_q(h, q));
    CK(cb_residual_sums(h, lpf));
    CK(cb_residual_allreduce(h));                 /* ncclAllReduce on the handle's stream, from C */
    double out[11], fused[11];
    CK(cb_get_sums(h, out));
    CK(cb_get_reaction_sums(h, out + 5));
    /* the same in one launch (sums + exchange over NVLink peer memory where the ranks could map each other) */
    CK(cb_residual_sums_allreduce(h, lpf));
    CK(cb_get_sums(h, fused));
    CK(cb_get_reaction_sums(h, fused + 5));
    if (memcmp(out, fused, sizeof out)) { fprintf(stderr, "rank %d: fused sums + all-reduce differ from the two calls\n", rank); return 1; }
    FILE *o = fopen(argv[5], "wb"); if (!o) { perror(argv[5]); return 2; }
    const double mode = cb_comm_peer_memory(h);
    fwrite(out, sizeof(double), 11, o); fwrite(&mode, sizeof mode, 1, o); fclose(o);
    const int peer = cb_comm_peer_memory(h);
    CK(cb_comm_destroy(h));
    cb_destroy(h);
    printf("rank %d of %d: sums %.17g %.17g %.17g (all-reduce over %s)\n", rank, world, out[0], out[1], out[2],
           peer ? "mapped peer memory" : "NCCL");
    return 0;
}
