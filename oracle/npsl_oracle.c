/* TEST INFRASTRUCTURE ONLY -- parity checker, never part of the product path.
 *
 * Plain-C restatement of np-shape-lab's per-step force / energy / integrator path over flat
 * arrays. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library. Pinned against the reference itself: tests/test_oracle.py checks
 * every function below against tests/golden/ (vectors produced by the reference's unmodified
 * sources, oracle/_ref, see oracle/gen_golden.py) and against the 6-digit series the reference
 * ships in examples/.
 *
 * Arithmetic follows the reference's types: 3-vectors are x87 long double
 * (src/vector3d.h:23), most scalars are truncated to double where the reference does so.
 * All citations are file:line in /root/reference.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef long double ld;
typedef struct { ld x, y, z; } vec3;

static inline vec3 v3(ld x, ld y, ld z) { vec3 r = {x, y, z}; return r; }
static inline vec3 vsub(vec3 a, vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline vec3 vadd(vec3 a, vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline vec3 vscale(vec3 a, ld s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline ld vdot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline vec3 vcross(vec3 a, vec3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline ld vnorm(vec3 a) { return sqrtl(a.x * a.x + a.y * a.y + a.z * a.z); }

#define NPO_MAXDEG 16
static const double dcut2 = 1.25992104989487316476; /* src/utility.h:37 */
static const double kB = 1.0;                        /* src/utility.h:49 */

typedef struct { double Q, T, dof, xi, eta; } bath_t; /* src/thermostat.h:28-35 */

typedef struct {
    int nv, ne, nf;
    int *edge_v;   /* [ne][2] */
    int *face_v;   /* [nf][3], stored winding */
    int *edge_f;   /* [ne][2], in the order faces register themselves, src/interface.cpp:134-142 */
    int *v_ne, *v_e; /* per vertex incident edges, file order (src/interface.cpp:108-109) */
    int *v_nf, *v_f; /* per vertex incident faces, file order (src/interface.cpp:130-132) */
    double *l0;
    /* parameters */
    double scalefactor, em, inv_kappa_out, elj, lj_length_mesh_mesh, bkappa, sconstant, sigma_a, sigma_v;
    double ref_area, ref_volume, avg_edge_length, dt, T;
    int buckling;
    /* state */
    vec3 *pos, *vel, *forvec, *pforce, *bforce, *sforce, *tforce, *vforce, *negGradV, *negGradA;
    double *q, *ke;
    vec3 *f_dir, *f_gradA; /* f_gradA[f*3+c] = areaGradient[(corner c of f, f)] */
    double *f_area, *f_vol, *e_S;
    double total_area, total_volume;
    /* energies, src/interface.h:77 */
    double netMeshKE, penergy, energy, bE, sE, tE, volE, ljE, esE;
    ld vertex_ke;
    int chain;
    bath_t mesh_bath[16], ions_bath[16];
    int constraint; /* 1: rigid volume constraint, geomConstraint 'V', linear form */
    /* explicit counterions, src/particle.h (mass 1); all of one diameter */
    int n_ions;
    vec3 *ion_pos, *ion_vel, *ion_for;
    double *ion_q, *ion_ke;
    double ion_diameter, lj_length_mesh_ions, netIonKE;
    ld ions_ke;
} npo_t;

/* ---------------------------------------------------------------- construction */

static int edge_between(const npo_t *h, int a, int b) { /* src/interface.cpp:312-319 */
    for (int k = 0; k < h->v_ne[a]; k++) {
        int e = h->v_e[a * NPO_MAXDEG + k];
        int o = (h->edge_v[2 * e] == a) ? h->edge_v[2 * e + 1] : (h->edge_v[2 * e + 1] == a ? h->edge_v[2 * e] : -1);
        if (o == b) return e;
    }
    return -1;
}

/* params: scalefactor em inv_kappa_out elj lj_length_mesh_mesh bkappa sconstant sigma_a sigma_v
 *         ref_area ref_volume avg_edge_length dt T buckling     (15 doubles) */
void *npo_create(int nv, int ne, int nf, const int *edges, const int *faces, const double *l0, const double *p) {
    npo_t *h = (npo_t *) calloc(1, sizeof(npo_t));
    h->nv = nv; h->ne = ne; h->nf = nf;
    h->edge_v = (int *) malloc(sizeof(int) * 2 * ne);
    h->face_v = (int *) malloc(sizeof(int) * 3 * nf);
    h->edge_f = (int *) malloc(sizeof(int) * 2 * ne);
    memcpy(h->edge_v, edges, sizeof(int) * 2 * ne);
    memcpy(h->face_v, faces, sizeof(int) * 3 * nf);
    h->v_ne = (int *) calloc(nv, sizeof(int));
    h->v_nf = (int *) calloc(nv, sizeof(int));
    h->v_e = (int *) malloc(sizeof(int) * nv * NPO_MAXDEG);
    h->v_f = (int *) malloc(sizeof(int) * nv * NPO_MAXDEG);
    h->l0 = (double *) malloc(sizeof(double) * ne);
    memcpy(h->l0, l0, sizeof(double) * ne);
    for (int e = 0; e < ne; e++) {
        for (int k = 0; k < 2; k++) {
            int v = edges[2 * e + k];
            if (h->v_ne[v] >= NPO_MAXDEG) { free(h); return NULL; }
            h->v_e[v * NPO_MAXDEG + h->v_ne[v]++] = e;
        }
        h->edge_f[2 * e] = h->edge_f[2 * e + 1] = -1;
    }
    for (int f = 0; f < nf; f++) {
        const int *fv = faces + 3 * f;
        /* the reference registers faces on the *file* columns; for the vertex's face list
         * only membership and file order of faces matter (src/interface.cpp:130-132) */
        for (int k = 0; k < 3; k++) {
            int v = fv[k];
            if (h->v_nf[v] >= NPO_MAXDEG) { free(h); return NULL; }
            h->v_f[v * NPO_MAXDEG + h->v_nf[v]++] = f;
        }
        int pairs[3][2] = {{fv[1], fv[2]}, {fv[2], fv[0]}, {fv[0], fv[1]}};
        for (int k = 0; k < 3; k++) {
            int e = edge_between(h, pairs[k][0], pairs[k][1]);
            if (e < 0) { free(h); return NULL; }
            if (h->edge_f[2 * e] < 0) h->edge_f[2 * e] = f;
            else h->edge_f[2 * e + 1] = f;
        }
    }
    h->scalefactor = p[0]; h->em = p[1]; h->inv_kappa_out = p[2]; h->elj = p[3];
    h->lj_length_mesh_mesh = p[4]; h->bkappa = p[5]; h->sconstant = p[6]; h->sigma_a = p[7];
    h->sigma_v = p[8]; h->ref_area = p[9]; h->ref_volume = p[10]; h->avg_edge_length = p[11];
    h->dt = p[12]; h->T = p[13]; h->buckling = (int) p[14];
#define ALLOCV(name, n) h->name = (vec3 *) calloc((n), sizeof(vec3))
    ALLOCV(pos, nv); ALLOCV(vel, nv); ALLOCV(forvec, nv); ALLOCV(pforce, nv); ALLOCV(bforce, nv);
    ALLOCV(sforce, nv); ALLOCV(tforce, nv); ALLOCV(vforce, nv); ALLOCV(negGradV, nv); ALLOCV(negGradA, nv);
    ALLOCV(f_dir, nf); ALLOCV(f_gradA, 3 * nf);
    h->q = (double *) calloc(nv, sizeof(double));
    h->ke = (double *) calloc(nv, sizeof(double));
    h->f_area = (double *) calloc(nf, sizeof(double));
    h->f_vol = (double *) calloc(nf, sizeof(double));
    h->e_S = (double *) calloc(ne, sizeof(double));
    return h;
}

void npo_destroy(void *hh) {
    npo_t *h = (npo_t *) hh;
    if (!h) return;
    free(h->edge_v); free(h->face_v); free(h->edge_f); free(h->v_ne); free(h->v_nf); free(h->v_e); free(h->v_f);
    free(h->l0); free(h->pos); free(h->vel); free(h->forvec); free(h->pforce); free(h->bforce); free(h->sforce);
    free(h->tforce); free(h->vforce); free(h->negGradV); free(h->negGradA); free(h->f_dir); free(h->f_gradA);
    free(h->q); free(h->ke); free(h->f_area); free(h->f_vol); free(h->e_S);
    free(h->ion_pos); free(h->ion_vel); free(h->ion_for); free(h->ion_q); free(h->ion_ke);
    free(h);
}

void npo_set_state(void *hh, const double *pos, const double *vel, const double *q) {
    npo_t *h = (npo_t *) hh;
    for (int i = 0; i < h->nv; i++) {
        if (pos) h->pos[i] = v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        if (vel) h->vel[i] = v3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        if (q) h->q[i] = q[i];
    }
}

/* baths: [chain][5] rows (Q, T, dof, xi, eta), src/main.cpp:276-290 */
void npo_set_baths(void *hh, int chain, const double *mesh, const double *ions) {
    npo_t *h = (npo_t *) hh;
    h->chain = chain;
    for (int j = 0; j < chain; j++) {
        bath_t m = {mesh[5 * j], mesh[5 * j + 1], mesh[5 * j + 2], mesh[5 * j + 3], mesh[5 * j + 4]};
        bath_t n = {ions[5 * j], ions[5 * j + 1], ions[5 * j + 2], ions[5 * j + 3], ions[5 * j + 4]};
        h->mesh_bath[j] = m;
        h->ions_bath[j] = n;
    }
}

void npo_get_baths(void *hh, double *mesh, double *ions) {
    npo_t *h = (npo_t *) hh;
    for (int j = 0; j < h->chain; j++) {
        bath_t m = h->mesh_bath[j], n = h->ions_bath[j];
        mesh[5 * j] = m.Q; mesh[5 * j + 1] = m.T; mesh[5 * j + 2] = m.dof; mesh[5 * j + 3] = m.xi; mesh[5 * j + 4] = m.eta;
        ions[5 * j] = n.Q; ions[5 * j + 1] = n.T; ions[5 * j + 2] = n.dof; ions[5 * j + 3] = n.xi; ions[5 * j + 4] = n.eta;
    }
}

/* ---------------------------------------------------------------- geometry */

/* INTERFACE::dressup, src/interface.cpp:33-67; FACE::compute_area_normal src/face.cpp:16-26;
 * FACE::computevolume src/face.cpp:38-45 */
void npo_dressup(void *hh) {
    npo_t *h = (npo_t *) hh;
    for (int f = 0; f < h->nf; f++) {
        vec3 a = h->pos[h->face_v[3 * f]], b = h->pos[h->face_v[3 * f + 1]], c = h->pos[h->face_v[3 * f + 2]];
        h->f_dir[f] = vcross(vsub(b, a), vsub(c, b));
        h->f_area[f] = (double) (0.5 * vnorm(h->f_dir[f]));
        h->f_vol[f] = (double) ((1. / 6.) * vdot(a, vcross(b, c)));
    }
    h->total_area = 0;
    for (int f = 0; f < h->nf; f++) h->total_area += h->f_area[f];
    h->total_volume = 0;
    for (int f = 0; f < h->nf; f++) h->total_volume += h->f_vol[f];
}

/* FACE::computegradients, src/face.cpp:47-91 (corner c of face f) */
static void face_gradients(const npo_t *h, int f, int c, vec3 *gradA, vec3 *gradV) {
    vec3 p = h->pos[h->face_v[3 * f + c]];
    vec3 n = h->pos[h->face_v[3 * f + (c + 1) % 3]];
    vec3 nn = h->pos[h->face_v[3 * f + (c + 2) % 3]];
    *gradV = v3((1. / 6.) * (n.y * nn.z - n.z * nn.y), (1. / 6.) * (n.z * nn.x - n.x * nn.z),
                (1. / 6.) * (n.x * nn.y - n.y * nn.x));
    vec3 cp = vcross(vsub(n, p), vsub(nn, p));
    double xc = (double) ((n.y - nn.y) * cp.z - (n.z - nn.z) * cp.y);
    double yc = (double) ((n.z - nn.z) * cp.x - (n.x - nn.x) * cp.z);
    double zc = (double) ((n.x - nn.x) * cp.y - (n.y - nn.y) * cp.x);
    double den = (double) (2 * vnorm(cp));
    *gradA = v3(xc / den, yc / den, zc / den);
}

/* VERTEX::compute_neg_VGrads for all vertices, src/vertex.cpp:34-53 */
static void neg_vgrads(npo_t *h) {
    for (int v = 0; v < h->nv; v++) {
        vec3 gv = v3(0, 0, 0), ga = v3(0, 0, 0);
        for (int k = 0; k < h->v_nf[v]; k++) {
            int f = h->v_f[v * NPO_MAXDEG + k];
            int c = (h->face_v[3 * f] == v) ? 0 : (h->face_v[3 * f + 1] == v ? 1 : 2);
            vec3 gA, gV;
            face_gradients(h, f, c, &gA, &gV);
            h->f_gradA[3 * f + c] = gA;
            gv = vsub(gv, gV);
            ga = vsub(ga, gA);
        }
        h->negGradV[v] = gv;
        h->negGradA[v] = ga;
    }
}

/* ---------------------------------------------------------------- pair term */

/* one LJ + ES pair force on the particle at p1 (charge q1) from the one at p2, LJ length d: the body shared by the
 * mesh-mesh, mesh-ion and ion-ion loops of src/newforces.cpp:153-250 */
static vec3 pair_force(const npo_t *h, vec3 p1, vec3 p2, double q1, double q2, double d) {
    vec3 lj = v3(0, 0, 0);
    vec3 r_vec = vsub(p1, p2);
    double r = (double) vnorm(r_vec);
    double r2 = (double) vdot(r_vec, r_vec);
    double d2 = d * d;
    if (r2 < dcut2 * d2) {
        double r6 = r2 * r2 * r2, r12 = r6 * r6, d6 = d2 * d2 * d2, d12 = d6 * d6;
        lj = vscale(r_vec, 48 * h->elj * ((d12 / r12) - 0.5 * (d6 / r6)) * (1 / r2));
    }
    double s1 = (-h->scalefactor * q1 * q2 / h->em) * (exp(-(1.0 / h->inv_kappa_out) * r));
    vec3 es = vscale(vscale(r_vec, (1 / r2)), s1);
    es = vscale(es, ((-1.0 / r) - (1.0 / h->inv_kappa_out)));
    return vadd(lj, es);
}

/* force_calculation mesh-mesh loop, src/newforces.cpp:153-177 (init twin :18-42), rows [lo,hi], followed by the
 * mesh-ion and ion-ion loops when there are counterions */
static void pair_forces(npo_t *h, int lo, int hi) {
    const int n = h->nv;
    for (int i = 0; i < n; i++) h->pforce[i] = v3(0, 0, 0);
#pragma omp parallel for schedule(dynamic)
    for (int i = lo; i <= hi; i++) {
        vec3 acc = v3(0, 0, 0);
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            vec3 lj = v3(0, 0, 0);
            vec3 r_vec = vsub(h->pos[i], h->pos[j]);
            double r = (double) vnorm(r_vec);
            double r2 = (double) vdot(r_vec, r_vec);
            double d2 = h->lj_length_mesh_mesh * h->lj_length_mesh_mesh;
            if (r2 < dcut2 * d2) {
                double r6 = r2 * r2 * r2, r12 = r6 * r6, d6 = d2 * d2 * d2, d12 = d6 * d6;
                lj = vscale(r_vec, 48 * h->elj * ((d12 / r12) - 0.5 * (d6 / r6)) * (1 / r2));
            }
            double s1 = (-h->scalefactor * h->q[i] * h->q[j] / h->em) * (exp(-(1.0 / h->inv_kappa_out) * r));
            vec3 es = vscale(vscale(r_vec, (1 / r2)), s1);
            es = vscale(es, ((-1.0 / r) - (1.0 / h->inv_kappa_out)));
            acc = vadd(acc, vadd(lj, es));
        }
        for (int j = 0; j < h->n_ions; j++) /* the mesh-ion component (for mesh), src/newforces.cpp:179-200 */
            acc = vadd(acc, pair_force(h, h->pos[i], h->ion_pos[j], h->q[i], h->ion_q[j], h->lj_length_mesh_ions));
        h->pforce[i] = acc;
    }
    /* ion rows, src/newforces.cpp:203-250 */
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < h->n_ions; i++) {
        vec3 acc = v3(0, 0, 0);
        for (int j = 0; j < h->n_ions; j++) {
            if (i == j) continue;
            acc = vadd(acc, pair_force(h, h->ion_pos[i], h->ion_pos[j], h->ion_q[i], h->ion_q[j], h->ion_diameter));
        }
        for (int j = 0; j < n; j++)
            acc = vadd(acc, pair_force(h, h->ion_pos[i], h->pos[j], h->ion_q[i], h->q[j], h->lj_length_mesh_ions));
        h->ion_for[i] = acc;
    }
}

static double pair_lj(const npo_t *h, vec3 p1, vec3 p2, double d) { /* energy_lj_pair, src/newenergies.cpp:4-17 */
    vec3 r_vec = vsub(p1, p2);
    double r2 = (double) vdot(r_vec, r_vec), d2 = d * d;
    if (r2 < dcut2 * d2) {
        double r6 = r2 * r2 * r2, d6 = d2 * d2 * d2;
        return 4 * h->elj * (d6 / r6) * ((d6 / r6) - 1) + h->elj;
    }
    return 0;
}
static double pair_es(const npo_t *h, vec3 p1, vec3 p2, double q1, double q2) { /* energy_es_pair, :21-40 */
    double r = (double) vnorm(vsub(p1, p2));
    return h->scalefactor * q1 * q2 * (exp(-(1.0 / h->inv_kappa_out) * r)) / (h->em * r);
}

/* energy_lj_pair / energy_es_pair, src/newenergies.cpp:4-25; loop src/interface.cpp:1046-1065 */
static void pair_energies(npo_t *h, int lo, int hi, double *lj_out, double *es_out) {
    const int n = h->nv;
    double *lj_row = (double *) calloc(n, sizeof(double)), *es_row = (double *) calloc(n, sizeof(double));
#pragma omp parallel for schedule(dynamic)
    for (int i = lo; i <= hi; i++) {
        double lj = 0, es = 0;
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            vec3 r_vec = vsub(h->pos[i], h->pos[j]);
            double r2 = (double) vdot(r_vec, r_vec);
            double d2 = h->lj_length_mesh_mesh * h->lj_length_mesh_mesh;
            if (r2 < dcut2 * d2) {
                double r6 = r2 * r2 * r2, d6 = d2 * d2 * d2;
                lj += 4 * h->elj * (d6 / r6) * ((d6 / r6) - 1) + h->elj;
            }
            double r = (double) vnorm(r_vec);
            es += h->scalefactor * h->q[i] * h->q[j] * (exp(-(1.0 / h->inv_kappa_out) * r)) / (h->em * r);
        }
        lj_row[i] = 0.5 * lj;
        es_row[i] = 0.5 * es;
        double lj_mi = 0, es_mi = 0; /* mesh-ion, counted once: src/interface.cpp:1059-1064 */
        for (int j = 0; j < h->n_ions; j++) {
            lj_mi += pair_lj(h, h->pos[i], h->ion_pos[j], h->lj_length_mesh_ions);
            es_mi += pair_es(h, h->pos[i], h->ion_pos[j], h->q[i], h->ion_q[j]);
        }
        lj_row[i] += lj_mi;
        es_row[i] += es_mi;
    }
    double lj_t = 0, es_t = 0;
    for (int i = lo; i <= hi; i++) { lj_t += lj_row[i]; es_t += es_row[i]; }
    for (int i = 0; i < h->n_ions; i++) { /* ion-ion, src/interface.cpp:1072-1089 */
        double lj = 0, es = 0;
        for (int j = 0; j < h->n_ions; j++) {
            if (i == j) continue;
            lj += pair_lj(h, h->ion_pos[i], h->ion_pos[j], h->ion_diameter);
            es += pair_es(h, h->ion_pos[i], h->ion_pos[j], h->ion_q[i], h->ion_q[j]);
        }
        lj_t += 0.5 * lj;
        es_t += 0.5 * es;
    }
    free(lj_row); free(es_row);
    *lj_out = lj_t; *es_out = es_t;
}

/* ---------------------------------------------------------------- mesh elastic terms */

static int across(const npo_t *h, int f, int e) { /* FACE::across(EDGE*), src/face.cpp:6-14 */
    int a = h->edge_v[2 * e], b = h->edge_v[2 * e + 1], r = -1;
    for (int k = 0; k < 3; k++) {
        int v = h->face_v[3 * f + k];
        if (v != a && v != b) r = v;
    }
    return r;
}

static int corner_of(const npo_t *h, int f, int v) {
    for (int k = 0; k < 3; k++) if (h->face_v[3 * f + k] == v) return k;
    return -1;
}

/* EDGE::compute_gradS, src/edge.cpp:60-152, restated through the identity
 *   grad_p S = sum over F in {F0,F1} containing p of (pn_F - pnn_F) x dir_{other(F)}
 * (pn, pnn follow p in F's stored winding); equal to the reference's expanded polynomial. */
static vec3 gradS(const npo_t *h, int e, int v) {
    vec3 g = v3(0, 0, 0);
    for (int s = 0; s < 2; s++) {
        int f = h->edge_f[2 * e + s], o = h->edge_f[2 * e + 1 - s];
        int c = corner_of(h, f, v);
        if (c < 0) continue;
        vec3 pn = h->pos[h->face_v[3 * f + (c + 1) % 3]], pnn = h->pos[h->face_v[3 * f + (c + 2) % 3]];
        g = vadd(g, vcross(vsub(pn, pnn), h->f_dir[o]));
    }
    return g;
}

/* EDGE::bending_forces for all edges, src/edge.cpp:38-58; src/newforces.cpp:254-257 */
static void bending_forces(npo_t *h) {
    for (int v = 0; v < h->nv; v++) h->bforce[v] = v3(0, 0, 0);
    for (int e = 0; e < h->ne; e++) {
        int f0 = h->edge_f[2 * e], f1 = h->edge_f[2 * e + 1];
        double S = (double) vdot(h->f_dir[f0], h->f_dir[f1]);
        h->e_S[e] = S;
        int vs[4] = {h->edge_v[2 * e], h->edge_v[2 * e + 1], across(h, f0, e), across(h, f1, e)};
        double a0 = h->f_area[f0], a1 = h->f_area[f1];
        for (int k = 0; k < 4; k++) {
            int v = vs[k];
            int c0 = corner_of(h, f0, v), c1 = corner_of(h, f1, v);
            vec3 g0 = c0 >= 0 ? h->f_gradA[3 * f0 + c0] : v3(0, 0, 0);
            vec3 g1 = c1 >= 0 ? h->f_gradA[3 * f1 + c1] : v3(0, 0, 0);
            vec3 t = vsub(gradS(h, e, v), vadd(vscale(g0, S / a0), vscale(g1, S / a1)));
            h->bforce[v] = vadd(h->bforce[v], vscale(t, h->bkappa * (0.25 / (a0 * a1))));
        }
    }
}

static double edge_length(const npo_t *h, int e) { /* src/edge.cpp:24-31 */
    vec3 d = vsub(h->pos[h->edge_v[2 * e + 1]], h->pos[h->edge_v[2 * e]]);
    return sqrt((double) vdot(d, d));
}

/* VERTEX::stretching_forces, src/vertex.cpp:56-89 */
static void stretching_forces(npo_t *h) {
    for (int v = 0; v < h->nv; v++) {
        vec3 s = v3(0, 0, 0);
        for (int k = 0; k < h->v_ne[v]; k++) {
            int e = h->v_e[v * NPO_MAXDEG + k];
            int o = (h->edge_v[2 * e] == v) ? h->edge_v[2 * e + 1] : h->edge_v[2 * e];
            double l = edge_length(h, e);
            if (!h->buckling) {
                double l0 = h->l0[e];
                vec3 t = vscale(vsub(h->pos[v], h->pos[o]), (l - l0) / l);
                s = vadd(s, vscale(t, 1.0 / (l0 * l0)));
            } else {
                double l0 = h->avg_edge_length;
                s = vadd(s, vscale(vsub(h->pos[v], h->pos[o]), (l - l0) / l));
            }
        }
        if (!h->buckling) s = vscale(s, (-1) * h->sconstant * h->avg_edge_length * h->avg_edge_length);
        else s = vscale(s, (-1) * h->sconstant);
        h->sforce[v] = s;
    }
}

/* VERTEX::tension_forces / volume_tension_forces, src/vertex.cpp:92-106 */
static void tension_forces(npo_t *h) {
    for (int v = 0; v < h->nv; v++) {
        h->tforce[v] = vscale(h->negGradA[v], h->sigma_a);
        h->vforce[v] = vscale(h->negGradV[v], 2.0 * h->sigma_v * (h->total_volume - h->ref_volume));
    }
}

/* force_calculation, src/newforces.cpp:137-289 (the init variant :4-135 sums the same terms).
 * Pair rows restricted to [lo,hi] as on one MPI rank; pass 0, nv-1 for the single-rank case.
 * Requires npo_dressup() for the current positions, as in the reference's loop (src/newmd.cpp:130). */
void npo_force(void *hh, int lo, int hi) {
    npo_t *h = (npo_t *) hh;
    neg_vgrads(h);
    pair_forces(h, lo, hi);
    bending_forces(h);
    stretching_forces(h);
    tension_forces(h);
    for (int v = 0; v < h->nv; v++)
        h->forvec[v] = vadd(vadd(vadd(vadd(h->pforce[v], h->bforce[v]), h->sforce[v]), h->tforce[v]), h->vforce[v]);
}

/* compute_kinetic_energy, src/newenergies.h:8-16; VERTEX::real_kinetic_energy src/vertex.h:111-114 */
static ld ion_kinetic_energy(npo_t *h) { /* src/newenergies.h:18-26, src/particle.h:86-90 */
    for (int i = 0; i < h->n_ions; i++) h->ion_ke[i] = (double) (0.5 * 1.0 * vdot(h->ion_vel[i], h->ion_vel[i]));
    ld ke = 0;
    for (int i = 0; i < h->n_ions; i++) ke += h->ion_ke[i];
    return ke;
}

static ld kinetic_energy(npo_t *h) {
    for (int v = 0; v < h->nv; v++) h->ke[v] = (double) (0.5 * 1.0 * vdot(h->vel[v], h->vel[v]));
    ld ke = 0;
    for (int v = 0; v < h->nv; v++) ke += h->ke[v];
    return ke;
}

/* INTERFACE::compute_energy, src/interface.cpp:964-1137. out: KE bE sE tE ljE esE volE penergy energy */
void npo_energy(void *hh, int lo, int hi, double *out) {
    npo_t *h = (npo_t *) hh;
    h->netMeshKE = 0;
    for (int v = 0; v < h->nv; v++) h->netMeshKE += h->ke[v];
    double b = 0;
    for (int e = 0; e < h->ne; e++)
        b += h->bkappa * (1 - h->e_S[e] / (4 * h->f_area[h->edge_f[2 * e]] * h->f_area[h->edge_f[2 * e + 1]]));
    double s = 0;
    if (!h->buckling) {
        for (int e = 0; e < h->ne; e++) {
            double st = edge_length(h, e) - h->l0[e];
            s += 0.5 * h->sconstant * st * st / (h->l0[e] * h->l0[e]);
        }
        s = s * h->avg_edge_length * h->avg_edge_length;
    } else {
        for (int e = 0; e < h->ne; e++) {
            double st = edge_length(h, e) - h->avg_edge_length;
            s += 0.5 * h->sconstant * st * st;
        }
    }
    h->bE = b; h->sE = s;
    h->tE = h->sigma_a * h->total_area;
    h->volE = h->sigma_v * ((h->total_volume - h->ref_volume) * (h->total_volume - h->ref_volume));
    pair_energies(h, lo, hi, &h->ljE, &h->esE);
    h->penergy = 0;
    h->penergy += h->bE; h->penergy += h->sE; h->penergy += h->tE; h->penergy += h->volE;
    h->penergy += h->ljE + h->esE;
    h->netIonKE = 0; /* src/interface.cpp:970-972 */
    for (int i = 0; i < h->n_ions; i++) h->netIonKE += h->ion_ke[i];
    h->energy = h->netMeshKE + h->netIonKE + h->penergy;
    out[0] = h->netMeshKE; out[1] = h->bE; out[2] = h->sE; out[3] = h->tE; out[4] = h->ljE; out[5] = h->esE;
    out[6] = h->volE; out[7] = h->penergy; out[8] = h->energy;
}

/* ---------------------------------------------------------------- data feeds of the writers */

/* VERTEX::compute_area_normal, src/vertex.cpp:7-18, after FACE::compute_area_normal (src/face.cpp:16-26) as the movie
 * writer refreshes them (src/newmd.cpp:232-235). area[nv], normal[nv][3]. Call after npo_dressup. */
void npo_vertex_geometry(void *hh, double *area, double *normal) {
    npo_t *h = (npo_t *) hh;
    for (int v = 0; v < h->nv; v++) {
        vec3 n = v3(0, 0, 0);
        for (int k = 0; k < h->v_nf[v]; k++) {
            int f = h->v_f[v * NPO_MAXDEG + k];
            vec3 fn = vscale(h->f_dir[f], 0.5 / h->f_area[f]); /* FACE::itsnormal */
            n = vadd(n, vscale(fn, h->f_area[f]));              /* area weighted */
        }
        n = vscale(n, 1 / vnorm(n));
        double a = 0;
        for (int k = 0; k < h->v_nf[v]; k++) a = a + h->f_area[h->v_f[v * NPO_MAXDEG + k]];
        a = a * (1. / 3.);
        area[v] = a;
        normal[3 * v] = (double) n.x; normal[3 * v + 1] = (double) n.y; normal[3 * v + 2] = (double) n.z;
    }
}

/* INTERFACE::compute_local_energies and compute_local_energies_by_component, src/interface.cpp:1140-1337: the
 * per-face values handed to colormap(). out[nf][4] = electrostatic, elastic, bending, stretching. Uses EDGE::itsS and
 * the face areas of the last npo_force / npo_dressup, like the reference. */
void npo_local_energies(void *hh, double *out) {
    npo_t *h = (npo_t *) hh;
    const int n = h->nv;
    double *es = (double *) calloc(h->nf, sizeof(double)), *el = (double *) calloc(h->nf, sizeof(double));
    double *be = (double *) calloc(h->nf, sizeof(double)), *st = (double *) calloc(h->nf, sizeof(double));
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            vec3 r_vec = vsub(h->pos[i], h->pos[j]);
            double r = (double) vnorm(r_vec);
            double cur_E = h->scalefactor * h->q[i] * h->q[j] * (exp(-(1.0 / h->inv_kappa_out) * r)) / (h->em * r);
            for (int k = 0; k < h->v_nf[i]; k++) es[h->v_f[i * NPO_MAXDEG + k]] += cur_E;
            for (int k = 0; k < h->v_nf[j]; k++) es[h->v_f[j * NPO_MAXDEG + k]] += cur_E;
        }
    for (int e = 0; e < h->ne; e++) {
        double stretched = edge_length(h, e) - h->l0[e];
        double senergy = 0.5 * h->sconstant * stretched * stretched / (h->l0[e] * h->l0[e]);
        senergy *= h->avg_edge_length * h->avg_edge_length;
        double benergy = h->bkappa * (1 - h->e_S[e] / (4 * h->f_area[h->edge_f[2 * e]] * h->f_area[h->edge_f[2 * e + 1]]));
        for (int k = 0; k < 2; k++) { el[h->edge_f[2 * e + k]] += senergy; st[h->edge_f[2 * e + k]] += senergy; }
        for (int k = 0; k < 2; k++) { el[h->edge_f[2 * e + k]] += benergy; be[h->edge_f[2 * e + k]] += benergy; }
    }
    for (int f = 0; f < h->nf; f++) {
        out[4 * f] = es[f]; out[4 * f + 1] = el[f]; out[4 * f + 2] = be[f]; out[4 * f + 3] = st[f];
    }
    free(es); free(el); free(be); free(st);
}

/* ---------------------------------------------------------------- thermostat + integrator */

/* update_chain_xi, src/functions.h:176-189 */
static void update_chain_xi(int j, bath_t *bath, double dt, ld ke) {
    if (bath[j].Q == 0) return;
    if (j != 0)
        bath[j].xi = bath[j].xi * exp(-0.5 * dt * bath[j + 1].xi) +
                     0.5 * dt * (1.0 / bath[j].Q) *
                         (bath[j - 1].Q * bath[j - 1].xi * bath[j - 1].xi - bath[j].dof * kB * bath[j].T) *
                         exp(-0.25 * dt * bath[j + 1].xi);
    else
        bath[j].xi = (double) (bath[j].xi * exp(-0.5 * dt * bath[j + 1].xi) +
                               0.5 * dt * (1.0 / bath[j].Q) * (2 * ke - bath[j].dof * kB * bath[j].T) *
                                   exp(-0.25 * dt * bath[j + 1].xi));
}

static void update_eta(bath_t *b, double dt) { /* src/thermostat.h:53-58 */
    if (b->Q == 0) return;
    b->eta = b->eta + 0.5 * dt * b->xi;
}

/* md_interface prologue, src/newmd.cpp:16-19,52-54: zero velocities, initial forces, KE */
void npo_md_init(void *hh) {
    npo_t *h = (npo_t *) hh;
    for (int v = 0; v < h->nv; v++) h->vel[v] = v3(0, 0, 0);
    for (int i = 0; i < h->n_ions; i++) h->ion_vel[i] = v3(0, 0, 0);
    npo_dressup(h);
    npo_force(h, 0, h->nv - 1);
    h->vertex_ke = kinetic_energy(h);
    h->ions_ke = ion_kinetic_energy(h);
}

/* compute_kinetic_energy at the current velocities (src/newmd.cpp:52) without the velocity reset of
 * the prologue: lets a test continue the loop body from an arbitrary (pos, vel). */
void npo_refresh_ke(void *hh) {
    npo_t *h = (npo_t *) hh;
    h->vertex_ke = kinetic_energy(h);
    h->ions_ke = ion_kinetic_energy(h);
}

/* gsl_poly_solve_cubic as the reference calls it (src/functions.h:102): GSL is a third-party dependency that is
 * not vendored in the reference tree (README.md:21 asks for GSL >= 2.x, unpinned); this restates the algorithm
 * its manual documents (Cardano / trigonometric solution, Numerical Recipes 5.6). x0 = smallest real root. */
static double cubic_x0(double a, double b, double c) {
    double q = a * a - 3 * b, r = 2 * a * a * a - 9 * a * b + 27 * c;
    double Q = q / 9, R = r / 54, Q3 = Q * Q * Q, R2 = R * R;
    double CR2 = 729 * r * r, CQ3 = 2916 * q * q * q;
    if (R == 0 && Q == 0) return -a / 3;
    if (CR2 == CQ3) { double sq = sqrt(Q); return (R > 0 ? -2 * sq : -sq) - a / 3; }
    if (R2 < Q3) {
        double sgnR = R >= 0 ? 1 : -1, theta = acos(sgnR * sqrt(R2 / Q3)), norm = -2 * sqrt(Q);
        double r0 = norm * cos(theta / 3) - a / 3, r1 = norm * cos((theta + 2.0 * M_PI) / 3) - a / 3,
               r2 = norm * cos((theta - 2.0 * M_PI) / 3) - a / 3;
        return fmin(r0, fmin(r1, r2));
    }
    double sgnR = R >= 0 ? 1 : -1, A = -sgnR * pow(fabs(R) + sqrt(R2 - Q3), 1.0 / 3.0);
    return A + Q / A - a / 3;
}

static ld triple(vec3 a, vec3 b, vec3 c) { return vdot(a, vcross(b, c)); }

/* SHAKE, src/functions.h:33-117 (linear form: neg_GradConi = neg_GradVi of the last force evaluation) */
static void shake(npo_t *h) {
    double f0 = 0, fa1 = 0, fa2 = 0, fa3 = 0, fb1 = 0, fb2 = 0, fb3 = 0, fc = 0;
    for (int f = 0; f < h->nf; f++) {
        const int *fv = h->face_v + 3 * f;
        vec3 p0 = h->pos[fv[0]], p1 = h->pos[fv[1]], p2 = h->pos[fv[2]];
        vec3 g0 = h->negGradV[fv[0]], g1 = h->negGradV[fv[1]], g2 = h->negGradV[fv[2]];
        f0 = f0 + triple(g0, g1, g2);
        fa1 = fa1 + triple(g0, p1, g2); fa2 = fa2 + triple(g0, g1, p2); fa3 = fa3 + triple(p0, g1, g2);
        fb1 = fb1 + triple(g0, p1, p2); fb2 = fb2 + triple(p0, p1, g2); fb3 = fb3 + triple(p0, g1, p2);
        fc = fc + triple(p0, p1, p2);
    }
    double idt = 1.0 / h->dt;
    double a = idt * (fa1 + fa2 + fa3) / f0, b = idt * idt * (fb1 + fb2 + fb3) / f0;
    double c = idt * idt * idt * (fc - 6 * h->ref_volume) / f0;
    if (c == 0 || f0 == 0) return;
    double x0 = cubic_x0(a, b, c);
    for (int v = 0; v < h->nv; v++) h->vel[v] = vadd(h->vel[v], vscale(h->negGradV[v], x0));
    for (int v = 0; v < h->nv; v++) h->pos[v] = vadd(h->pos[v], vscale(h->negGradV[v], x0 * h->dt));
}

/* RATTLE, src/functions.h:120-171 */
static void rattle(npo_t *h) {
    double num = 0, den = 0;
    for (int f = 0; f < h->nf; f++) {
        const int *fv = h->face_v + 3 * f;
        vec3 p0 = h->pos[fv[0]], p1 = h->pos[fv[1]], p2 = h->pos[fv[2]];
        vec3 v0 = h->vel[fv[0]], v1 = h->vel[fv[1]], v2 = h->vel[fv[2]];
        vec3 g0 = h->negGradV[fv[0]], g1 = h->negGradV[fv[1]], g2 = h->negGradV[fv[2]];
        num = num + triple(v0, p1, p2) + triple(p0, p1, v2) + triple(p0, v1, p2);
        den = den + triple(g0, p1, p2) + triple(p0, p1, g2) + triple(p0, g1, p2);
    }
    if (num == 0 || den == 0) return;
    double mu = -(num / den);
    for (int v = 0; v < h->nv; v++) h->vel[v] = vadd(h->vel[v], vscale(h->negGradV[v], mu));
}

void npo_set_constraint(void *hh, int on) { ((npo_t *) hh)->constraint = on; }

/* SHAKE + RATTLE prior to MD, src/newmd.cpp:33-38; call after npo_md_init */
void npo_constraint_init(void *hh) {
    npo_t *h = (npo_t *) hh;
    shake(h);
    rattle(h);
    h->vertex_ke = kinetic_energy(h);
}

/* One iteration of the loop of md_interface, src/newmd.cpp:96-170 (no ions; volume constraint optional) */
void npo_md_step(void *hh) {
    npo_t *h = (npo_t *) hh;
    const double dt = h->dt;
    const int C = h->chain;
    ld ions_ke = h->ions_ke;
    for (int j = C - 1; j > -1; j--) update_chain_xi(j, h->mesh_bath, dt, h->vertex_ke);
    for (int j = C - 1; j > -1; j--) update_chain_xi(j, h->ions_bath, dt, ions_ke);
    for (int j = 0; j < C; j++) update_eta(&h->mesh_bath[j], dt);
    for (int j = 0; j < C; j++) update_eta(&h->ions_bath[j], dt);
    double expfac = exp(-0.5 * dt * h->mesh_bath[0].xi);
    ld ef = expfac;
    ld kick = 0.5 * dt * sqrtl(ef); /* VERTEX::update_real_velocity, src/vertex.h:105-108 */
    ld ef_i = (ld) exp(-0.5 * dt * h->ions_bath[0].xi); /* expfac_ions is a long double holding a double, src/newmd.cpp:110 */
    ld kick_i = 0.5 * dt * sqrtl(ef_i);                  /* PARTICLE::update_real_velocity, src/particle.h:79-83 */
    for (int v = 0; v < h->nv; v++) h->vel[v] = vadd(vscale(h->vel[v], ef), vscale(h->forvec[v], kick));
    for (int i = 0; i < h->n_ions; i++) h->ion_vel[i] = vadd(vscale(h->ion_vel[i], ef_i), vscale(h->ion_for[i], kick_i));
    for (int v = 0; v < h->nv; v++) h->pos[v] = vadd(h->pos[v], vscale(h->vel[v], dt));
    for (int i = 0; i < h->n_ions; i++) h->ion_pos[i] = vadd(h->ion_pos[i], vscale(h->ion_vel[i], dt));
    if (h->constraint) shake(h); /* src/newmd.cpp:126 */
    npo_dressup(h);
    npo_force(h, 0, h->nv - 1);
    for (int v = 0; v < h->nv; v++) h->vel[v] = vadd(vscale(h->vel[v], ef), vscale(h->forvec[v], kick));
    for (int i = 0; i < h->n_ions; i++) h->ion_vel[i] = vadd(vscale(h->ion_vel[i], ef_i), vscale(h->ion_for[i], kick_i));
    if (h->constraint) rattle(h); /* src/newmd.cpp:152 */
    h->vertex_ke = kinetic_energy(h);
    h->ions_ke = ions_ke = ion_kinetic_energy(h);
    for (int j = 0; j < C; j++) update_eta(&h->mesh_bath[j], dt);
    for (int j = 0; j < C; j++) update_eta(&h->ions_bath[j], dt);
    for (int j = 0; j < C; j++) update_chain_xi(j, h->mesh_bath, dt, h->vertex_ke);
    for (int j = 0; j < C; j++) update_chain_xi(j, h->ions_bath, dt, ions_ke);
}

void npo_md_steps(void *hh, int n) {
    for (int k = 0; k < n; k++) npo_md_step(hh);
}

/* bath_kinetic_energy / bath_potential_energy, src/newenergies.h:28-46; src/thermostat.h:61-70
 * out: mesh_ke mesh_pe ions_ke ions_pe */
void npo_bath_energies(void *hh, double *out) {
    npo_t *h = (npo_t *) hh;
    double mk = 0, mp = 0, ik = 0, ip = 0;
    for (int j = 0; j < h->chain; j++) {
        mk += 0.5 * h->mesh_bath[j].Q * h->mesh_bath[j].xi * h->mesh_bath[j].xi;
        mp += h->mesh_bath[j].dof * kB * h->mesh_bath[j].T * h->mesh_bath[j].eta;
        ik += 0.5 * h->ions_bath[j].Q * h->ions_bath[j].xi * h->ions_bath[j].xi;
        ip += h->ions_bath[j].dof * kB * h->ions_bath[j].T * h->ions_bath[j].eta;
    }
    out[0] = mk; out[1] = mp; out[2] = ik; out[3] = ip;
}

/* ---------------------------------------------------------------- explicit counterions */

/* counterions as main() hands them to md_interface (vector<PARTICLE>, src/particle.h): positions, velocities, charges;
 * one diameter (the ion-ion LJ length, src/newforces.cpp:214) and INTERFACE::lj_length_mesh_ions (src/main.cpp:271-272) */
void npo_set_ions(void *hh, int n, const double *pos, const double *vel, const double *q, double diameter,
                  double lj_length_mesh_ions) {
    npo_t *h = (npo_t *) hh;
    free(h->ion_pos); free(h->ion_vel); free(h->ion_for); free(h->ion_q); free(h->ion_ke);
    h->n_ions = n;
    h->ion_pos = (vec3 *) calloc(n > 0 ? n : 1, sizeof(vec3));
    h->ion_vel = (vec3 *) calloc(n > 0 ? n : 1, sizeof(vec3));
    h->ion_for = (vec3 *) calloc(n > 0 ? n : 1, sizeof(vec3));
    h->ion_q = (double *) calloc(n > 0 ? n : 1, sizeof(double));
    h->ion_ke = (double *) calloc(n > 0 ? n : 1, sizeof(double));
    for (int i = 0; i < n; i++) {
        h->ion_pos[i] = v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        if (vel) h->ion_vel[i] = v3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        h->ion_q[i] = q[i];
    }
    h->ion_diameter = diameter;
    h->lj_length_mesh_ions = lj_length_mesh_ions;
    h->ions_ke = ion_kinetic_energy(h);
}

/* what: 0 pos 1 vel 2 force */
void npo_get_ions(void *hh, int what, double *out) {
    npo_t *h = (npo_t *) hh;
    const vec3 *src[3] = {h->ion_pos, h->ion_vel, h->ion_for};
    for (int i = 0; i < h->n_ions; i++) {
        out[3 * i] = (double) src[what][i].x; out[3 * i + 1] = (double) src[what][i].y; out[3 * i + 2] = (double) src[what][i].z;
    }
}

double npo_ions_ke(void *hh) { return (double) ((npo_t *) hh)->netIonKE; }

/* ---------------------------------------------------------------- accessors */

static void put(const vec3 *a, int n, double *out) {
    for (int i = 0; i < n; i++) { out[3 * i] = (double) a[i].x; out[3 * i + 1] = (double) a[i].y; out[3 * i + 2] = (double) a[i].z; }
}

/* what: 0 pos 1 vel 2 force 3 pair 4 bend 5 stretch 6 tension 7 voltension */
void npo_get_vec(void *hh, int what, double *out) {
    npo_t *h = (npo_t *) hh;
    const vec3 *src[8] = {h->pos, h->vel, h->forvec, h->pforce, h->bforce, h->sforce, h->tforce, h->vforce};
    put(src[what], h->nv, out);
}

/* out: total_area total_volume vertex_ke */
void npo_get_scalars(void *hh, double *out) {
    npo_t *h = (npo_t *) hh;
    out[0] = h->total_area; out[1] = h->total_volume; out[2] = (double) h->vertex_ke;
}

void npo_get_edge_S(void *hh, double *out) {
    npo_t *h = (npo_t *) hh;
    memcpy(out, h->e_S, sizeof(double) * h->ne);
}
