// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from the product path.
//
// Driver for the reference's own, unmodified translation units (compiled in place from
// /root/reference/src by oracle/Makefile into oracle/_ref/). It stands in for src/main.cpp,
// which cannot be built here (boost::program_options / boost::filesystem are absent), and
// performs the same set-up sequence before handing over to the reference's md_interface /
// force_calculation / INTERFACE::compute_energy:
//
//   unit scalings            src/main.cpp:196-210
//   constants                src/main.cpp:212-230
//   set_up + Debye length    src/main.cpp:231-239
//   discretize + charges     src/main.cpp:242-248
//   dressup + reference A,V  src/main.cpp:262-265
//   LJ parameters            src/main.cpp:268-273
//   thermostat chains        src/main.cpp:276-290
//   masses                   src/main.cpp:293-294
//   MPI row partition        src/main.cpp:448-465
//   md_interface             src/main.cpp:467
//
// Modes
//   --mode dump  : run md_interface for --steps K (K=0 -> only force_calculation_init and
//                  compute_energy(0)), then print the full state with 21 significant digits
//                  (long double) to --out. Used to generate tests/golden/.
//   --mode run   : run md_interface for --steps K exactly as the CLI would (series files land
//                  in <cwd>/outfiles), print wall time. Used as CPU baseline for small meshes.
//   --mode force : time --reps calls of force_calculation restricted to mesh rows
//                  [--row-lo, --row-hi] (the reference's own MPI row slice, src/main.cpp:448-455),
//                  print JSON. Used as CPU baseline for the synthetic large meshes.
//   --mode window: pins the large synthetic meshes (S64 / S128), where whole reference steps are too slow: optional
//                  deterministic displacement of every vertex (--perturb amp, fixed LCG), dressup, then ONE call of the
//                  reference's force_calculation per row window of --windows lo:hi,lo:hi,... (inclusive bounds, the
//                  reference's own lowerBoundMesh / upperBoundMesh, src/main.cpp:448-455, src/newforces.cpp:153-177).
//                  Writes raw float64 arrays into the directory --out: pos, q, the pair force of the window rows, and
//                  the mesh terms (bforce, sforce, TForce, VolTForce), EDGE::itsS, face areas / volumes in full.
//
// Run from a working directory that holds infiles/3_<D>.dat and an (empty) outfiles/ directory.
#include "utility.h"
#include "interface.h"
#include "vertex.h"
#include "edge.h"
#include "face.h"
#include "control.h"
#include "functions.h"
#include "thermostat.h"
#include "newforces.h"
#include "newenergies.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>

void md_interface(INTERFACE &, vector<PARTICLE> &, vector<THERMOSTAT> &, vector<THERMOSTAT> &, CONTROL &, char, char,
                  char, const double scalefactor, double box_halflength);

// globals the reference declares extern in src/mpi_utility.h:7-16 and defines in src/main.cpp:64-74
unsigned int lowerBoundMesh;
unsigned int upperBoundMesh;
unsigned int sizFVecMesh;
int lowerBoundIons;
int upperBoundIons;
int sizFVecIons;
unsigned int extraElementsIons;
mpi::environment env;
mpi::communicator world;

namespace {

struct Options {
    std::string mode = "dump";
    std::string out = "dump.txt";
    double R = 10, q = 600, c = 0.005, t = 1, v = 1, b = 40, s = 40;
    char B = 'n', F = 'n', H = 'n', G = 'N';
    int N = 1;
    double p = 0.5;
    int steps = 0;
    double dt = 0.001, T = 0.001, Qm = 0.01, Qi = 0.01;
    unsigned D = 4, C = 5;
    int writedata = 500;
    long row_lo = 0, row_hi = -1;
    int reps = 1;
    int warm = 0;
    int ions = 0;              // explicit counterions placed by place_ions() (the reference's own put_counterions cannot
                               // terminate with its hard-wired box_halflength = 1, src/main.cpp:194, src/interface.cpp:907-958)
    double alpha = 1.0;        // fractional charge occupancy (hard-wired to 1 in src/main.cpp:213)
    std::string E = "None";    // external charge pattern, infiles/Charge_Patterns/<E>.dat (src/main.cpp:244-247)
    int local = 0;  // 1: also dump vertex areas / normals and the per-face local energy maps
    std::string windows;       // --mode window: "lo:hi,lo:hi,..." inclusive row bounds
    double perturb = 0;        // --mode window: amplitude of the deterministic displacement (0: none)
    int anneal_freq = 50000, anneal_duration = 5000;  // CONTROL::annealfreq / annealDuration (src/main.cpp:150-153)
};

bool parse(int argc, char **argv, Options &o) {
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](const char *name) -> const char * {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "missing value for %s\n", name);
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--mode") o.mode = need("--mode");
        else if (a == "--out") o.out = need("--out");
        else if (a == "-R") o.R = atof(need("-R"));
        else if (a == "-q") o.q = atof(need("-q"));
        else if (a == "-c") o.c = atof(need("-c"));
        else if (a == "-t") o.t = atof(need("-t"));
        else if (a == "-v") o.v = atof(need("-v"));
        else if (a == "-b") o.b = atof(need("-b"));
        else if (a == "-s") o.s = atof(need("-s"));
        else if (a == "-B") o.B = need("-B")[0];
        else if (a == "-F") o.F = need("-F")[0];
        else if (a == "-H") o.H = need("-H")[0];
        else if (a == "-G") o.G = need("-G")[0];
        else if (a == "-N") o.N = atoi(need("-N"));
        else if (a == "-p") o.p = atof(need("-p"));
        else if (a == "-S" || a == "--steps") o.steps = atoi(need("-S"));
        else if (a == "-d") o.dt = atof(need("-d"));
        else if (a == "-T") o.T = atof(need("-T"));
        else if (a == "-Q") o.Qm = atof(need("-Q"));
        else if (a == "-Y") o.Qi = atof(need("-Y"));
        else if (a == "-D") o.D = (unsigned) atoi(need("-D"));
        else if (a == "-C") o.C = (unsigned) atoi(need("-C"));
        else if (a == "-W") o.writedata = atoi(need("-W"));
        else if (a == "--row-lo") o.row_lo = atol(need("--row-lo"));
        else if (a == "--row-hi") o.row_hi = atol(need("--row-hi"));
        else if (a == "--reps") o.reps = atoi(need("--reps"));
        else if (a == "--warm") o.warm = atoi(need("--warm"));
        else if (a == "--local") o.local = atoi(need("--local"));
        else if (a == "--alpha") o.alpha = atof(need("--alpha"));
        else if (a == "--ions") o.ions = atoi(need("--ions"));
        else if (a == "-E") o.E = need("-E");
        else if (a == "--windows") o.windows = need("--windows");
        else if (a == "--perturb") o.perturb = atof(need("--perturb"));
        else if (a == "--anneal-freq") o.anneal_freq = atoi(need("--anneal-freq"));
        else if (a == "--anneal-duration") o.anneal_duration = atoi(need("--anneal-duration"));
        else {
            std::fprintf(stderr, "unknown option %s\n", a.c_str());
            return false;
        }
    }
    return true;
}

void pv(FILE *f, const char *name, size_t i, VECTOR3D &a) {
    std::fprintf(f, "%s %zu %.21Lg %.21Lg %.21Lg\n", name, i, a.x, a.y, a.z);
}

// Deterministic stand-in for INTERFACE::put_counterions: monovalent counterions (PARTICLE(id, diameter, valency,
// -valency, m = 1, pos), src/interface.cpp:944) on a shell around the unit sphere from a fixed LCG; a few are put close
// to the membrane (inside the mesh-ion WCA cutoff) and two closer to each other than one ion diameter, so that every
// branch of the mesh-ion and ion-ion loops (src/newforces.cpp:44-65,179-250) is exercised.
void place_ions(std::vector<PARTICLE> &ions, int n, double diameter) {
    unsigned long long st = 88172645463325252ULL;
    auto rnd = [&]() {
        st = st * 6364136223846793005ULL + 1442695040888963407ULL;
        return (double) (st >> 11) * (1.0 / 9007199254740992.0);
    };
    for (int k = 0; k < n; k++) {
        double z = 2 * rnd() - 1, phi = 6.283185307179586 * rnd();
        double rad = (k % 16 == 0) ? 1.09 + 0.03 * rnd() : 1.3 + 0.5 * rnd();
        double s = std::sqrt(1 - z * z);
        VECTOR3D p(rad * s * std::cos(phi), rad * s * std::sin(phi), rad * z);
        if (k == 1 && n > 1) p = ions[0].posvec + VECTOR3D(0.8 * diameter, 0.1 * diameter, 0);  // overlapping pair
        ions.push_back(PARTICLE(k + 1, diameter, 1, -1.0, 1.0, p));
    }
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

int main(int argc, char **argv) {
    Options o;
    if (!parse(argc, argv, o)) return 2;

    INTERFACE boundary;
    CONTROL mdremote;
    vector<PARTICLE> counterions;  // empty unless --ions M

    boundary.sigma_a = o.t;
    boundary.sigma_v = o.v;
    boundary.bkappa = o.b;
    boundary.sconstant = o.s;
    mdremote.steps = o.steps;
    mdremote.timestep = o.dt;
    mdremote.anneal = 'y';
    mdremote.annealfreq = o.anneal_freq;
    mdremote.annealDuration = o.anneal_duration;
    mdremote.TAnnealFac = 1.25;
    mdremote.moviestart = 1;
    mdremote.offfreq = 2500;
    mdremote.moviefreq = 1000;
    mdremote.povfreq = 100000;
    mdremote.writedata = o.writedata;

    const double unit_radius_sphere = o.R;
    double box_halflength = 1.0;
    double counterion_diameter = 0.6 / unit_radius_sphere;

    double dynePerCm_Scalefactor = pow(10, -14) * (pow(unit_radius_sphere, 2) / unitenergy);
    boundary.sigma_a = (dynePerCm_Scalefactor * boundary.sigma_a);
    double atmPerNmThird_Scalefactor =
        (1.01325 * pow(10, 5)) * pow(10, -27) * (pow(unit_radius_sphere, 6) / (unitenergy * pow(10, -7)));
    boundary.sigma_v = atmPerNmThird_Scalefactor * boundary.sigma_v;
    const double scalefactor = epsilon_water * lB_water / unit_radius_sphere;

    unsigned int chain_length = o.C;
    double Q_mesh = o.Qm, Q_ion = o.Qi, T = o.T;
    if (chain_length == 1) Q_mesh = 0;
    if (chain_length == 1) Q_ion = 0;
    mdremote.QAnnealFac = 1.00 + (mdremote.TAnnealFac / 10.0);
    char constraintForm = 'L';
    double lambda_v = 0, lambda_a = 0;
    boundary.lambda_l = 0;
    double z_out = 1;
    boundary.ein = epsilon_water;
    boundary.eout = epsilon_water;
    boundary.set_up(unit_radius_sphere);
    if (o.c == 0)
        boundary.inv_kappa_out = 100000000;
    else
        boundary.inv_kappa_out = (0.257 / sqrt(z_out * boundary.lB_out * unit_radius_sphere * o.c)) / unit_radius_sphere;

    boundary.discretize(3, o.D);
    if (o.E == "None")
        boundary.assign_random_q_values(o.q, o.alpha, o.N, o.p, o.F, o.H);
    else
        boundary.assign_external_q_values(o.q, o.E);
    boundary.dressup(lambda_a, lambda_v);
    boundary.ref_area = boundary.total_area;
    boundary.ref_volume = boundary.total_volume;

    boundary.lj_length_mesh_mesh = 0.1 * boundary.avg_edge_length;
    double meshPointRadius = 0.5 * boundary.avg_edge_length;
    boundary.lj_length_mesh_ions = ((2.0 * meshPointRadius + counterion_diameter) / 2.0);
    boundary.elj = 1.0;

    if (o.ions > 0) place_ions(counterions, o.ions, counterion_diameter);
    vector<THERMOSTAT> mesh_bath, ions_bath;
    if (chain_length == 1) {
        mesh_bath.push_back(THERMOSTAT(0, T, 3 * boundary.V.size(), 0.0, 0, 0));
        ions_bath.push_back(THERMOSTAT(0, T, 3 * counterions.size(), 0.0, 0, 0));
    } else {
        mesh_bath.push_back(THERMOSTAT(Q_mesh, T, 3 * boundary.V.size(), 0, 0, 0));
        ions_bath.push_back(THERMOSTAT(Q_ion, T, 3 * counterions.size(), 0, 0, 0));
        while (mesh_bath.size() != chain_length - 1) mesh_bath.push_back(THERMOSTAT(Q_mesh, T, 1, 0, 0, 0));
        while (ions_bath.size() != chain_length - 1) ions_bath.push_back(THERMOSTAT(Q_ion, T, 1, 0, 0, 0));
        mesh_bath.push_back(THERMOSTAT(0, T, 3 * boundary.V.size(), 0.0, 0, 0));
        ions_bath.push_back(THERMOSTAT(0, T, 3 * counterions.size(), 0.0, 0, 0));
    }
    for (unsigned int i = 0; i < boundary.V.size(); i++) boundary.V[i].m = 1.0;
    boundary.dressup(lambda_a, lambda_v);

    // single-rank partition of src/main.cpp:448-465
    lowerBoundMesh = 0;
    upperBoundMesh = boundary.V.size() - 1;
    sizFVecMesh = boundary.V.size();
    lowerBoundIons = 0;
    upperBoundIons = (int) counterions.size() - 1;
    sizFVecIons = counterions.size();

    const size_t N = boundary.V.size();

    if (o.mode == "force") {
        // One rank's row slice of the reference's decomposition; everything else in
        // force_calculation (mesh terms) is evaluated in full, as on every MPI rank.
        long lo = o.row_lo, hi = (o.row_hi < 0 ? (long) N - 1 : o.row_hi);
        if (hi > (long) N - 1) hi = (long) N - 1;
        lowerBoundMesh = (unsigned) lo;
        upperBoundMesh = (unsigned) hi;
        sizFVecMesh = (unsigned) (hi - lo + 1);
        initialize_velocities_to_zero(boundary.V, counterions);
        double best = 1e300, total = 0;
        for (int r = 0; r < o.warm; r++) force_calculation(boundary, counterions, scalefactor, o.B);  // untimed
        for (int r = 0; r < o.reps; r++) {
            double t0 = now_s();
            force_calculation(boundary, counterions, scalefactor, o.B);
            double t1 = now_s();
            total += t1 - t0;
            if (t1 - t0 < best) best = t1 - t0;
        }
        int threads = 1;
#pragma omp parallel
        {
#pragma omp master
            threads = omp_get_num_threads();
        }
        double pairs = (double) (hi - lo + 1) * (double) (N - 1);
        std::printf("{\"mode\": \"force\", \"vertices\": %zu, \"rows\": %ld, \"pairs_per_call\": %.0f, \"reps\": %d, "
                    "\"best_s\": %.6f, \"mean_s\": %.6f, \"threads\": %d, \"fx0\": %.17Lg}\n",
                    N, hi - lo + 1, pairs, o.reps, best, total / o.reps, threads, boundary.V[lo].forvec.x);
        return 0;
    }

    if (o.mode == "window") {
        initialize_velocities_to_zero(boundary.V, counterions);
        if (o.perturb > 0) {  // every coordinate moves by perturb * (2u - 1), u from a fixed 64-bit LCG (top 53 bits)
            unsigned long long st = 0x9E3779B97F4A7C15ULL;
            for (size_t i = 0; i < N; i++) {
                double d[3];
                for (int k = 0; k < 3; k++) {
                    st = st * 6364136223846793005ULL + 1442695040888963407ULL;
                    d[k] = o.perturb * (2.0 * ((double) (st >> 11) * (1.0 / 9007199254740992.0)) - 1.0);
                }
                // added in double, so that the displaced coordinates are exactly representable on the device side too
                const double nx = (double) boundary.V[i].posvec.x + d[0], ny = (double) boundary.V[i].posvec.y + d[1],
                             nz = (double) boundary.V[i].posvec.z + d[2];
                boundary.V[i].posvec = VECTOR3D(nx, ny, nz);
            }
        }
        boundary.dressup(lambda_a, lambda_v);  // src/newmd.cpp:130, before force_calculation
        auto put = [&](const std::string &name, const std::vector<double> &v) {
            FILE *g = std::fopen((o.out + "/" + name + ".f64").c_str(), "wb");
            if (!g) { std::perror("open output"); std::exit(1); }
            std::fwrite(v.data(), sizeof(double), v.size(), g);
            std::fclose(g);
        };
        auto vec3 = [&](std::vector<double> &v, VECTOR3D a) {
            v.push_back((double) a.x); v.push_back((double) a.y); v.push_back((double) a.z);
        };
        std::vector<double> rows, pfo;
        std::string w = o.windows;
        double t_force = 0;
        while (!w.empty()) {
            size_t comma = w.find(',');
            std::string one = w.substr(0, comma);
            w = comma == std::string::npos ? "" : w.substr(comma + 1);
            long lo = atol(one.c_str()), hi = atol(one.substr(one.find(':') + 1).c_str());
            if (hi > (long) N - 1) hi = (long) N - 1;
            lowerBoundMesh = (unsigned) lo;
            upperBoundMesh = (unsigned) hi;
            sizFVecMesh = (unsigned) (hi - lo + 1);
            double t0 = now_s();
            force_calculation(boundary, counterions, scalefactor, o.B);
            t_force += now_s() - t0;
            for (long i = lo; i <= hi; i++) {
                VERTEX &V = boundary.V[i];
                rows.push_back((double) i);
                vec3(pfo, V.forvec - V.bforce - V.sforce - V.TForce - V.VolTForce);
            }
        }
        std::vector<double> pos, q, bfo, sfo, tfo, vfo, itsS, farea, fvol;
        for (size_t i = 0; i < N; i++) {
            VERTEX &V = boundary.V[i];
            vec3(pos, V.posvec); q.push_back(V.q);
            vec3(bfo, V.bforce); vec3(sfo, V.sforce); vec3(tfo, V.TForce); vec3(vfo, V.VolTForce);
        }
        for (size_t i = 0; i < boundary.E.size(); i++) itsS.push_back(boundary.E[i].itsS);
        for (size_t i = 0; i < boundary.F.size(); i++) {
            farea.push_back(boundary.F[i].itsarea);
            fvol.push_back(boundary.F[i].subtends_volume);
        }
        put("rows", rows); put("pforce", pfo); put("pos", pos); put("q", q);
        put("bforce", bfo); put("sforce", sfo); put("tforce", tfo); put("vforce", vfo);
        put("itsS", itsS); put("face_area", farea); put("face_volume", fvol);
        std::printf("{\"mode\": \"window\", \"vertices\": %zu, \"edges\": %zu, \"faces\": %zu, \"rows\": %zu, "
                    "\"force_s\": %.3f, \"scalefactor\": %.17g, \"em\": %.17g, \"inv_kappa_out\": %.17g, \"elj\": %.17g, "
                    "\"lj_length_mesh_mesh\": %.17g, \"bkappa\": %.17g, \"sconstant\": %.17g, \"sigma_a\": %.17g, "
                    "\"sigma_v\": %.17g, \"ref_area\": %.17g, \"ref_volume\": %.17g, \"avg_edge_length\": %.17g, "
                    "\"total_area\": %.17g, \"total_volume\": %.17g}\n",
                    N, boundary.E.size(), boundary.F.size(), rows.size(), t_force, scalefactor, boundary.em,
                    boundary.inv_kappa_out, boundary.elj, boundary.lj_length_mesh_mesh, boundary.bkappa, boundary.sconstant,
                    boundary.sigma_a, boundary.sigma_v, boundary.ref_area, boundary.ref_volume, boundary.avg_edge_length,
                    boundary.total_area, boundary.total_volume);
        return 0;
    }

    double t0 = now_s();
    md_interface(boundary, counterions, mesh_bath, ions_bath, mdremote, o.G, o.B, constraintForm, scalefactor,
                 box_halflength);
    double t1 = now_s();

    if (o.mode == "run") {
        int threads = 1;
#pragma omp parallel
        {
#pragma omp master
            threads = omp_get_num_threads();
        }
        std::printf("{\"mode\": \"run\", \"vertices\": %zu, \"steps\": %d, \"wall_s\": %.6f, \"threads\": %d, "
                    "\"area\": %.17g, \"penergy\": %.17g, \"energy\": %.17g}\n",
                    N, o.steps, t1 - t0, threads, boundary.total_area, boundary.penergy, boundary.energy);
        return 0;
    }

    // ---- dump ----
    FILE *f = std::fopen(o.out.c_str(), "w");
    if (!f) {
        std::perror("open --out");
        return 1;
    }
    std::fprintf(f, "sizes %zu %zu %zu\n", N, boundary.E.size(), boundary.F.size());
    std::fprintf(f, "step %d\n", o.steps);
    std::fprintf(f, "scalefactor %.17g\nem %.17g\ninv_kappa_out %.17g\nelj %.17g\nlj_length_mesh_mesh %.17g\n",
                 scalefactor, boundary.em, boundary.inv_kappa_out, boundary.elj, boundary.lj_length_mesh_mesh);
    std::fprintf(f, "bkappa %.17g\nsconstant %.17g\nsigma_a %.17g\nsigma_v %.17g\n", boundary.bkappa,
                 boundary.sconstant, boundary.sigma_a, boundary.sigma_v);
    std::fprintf(f, "ref_area %.17g\nref_volume %.17g\navg_edge_length %.17g\n", boundary.ref_area,
                 boundary.ref_volume, boundary.avg_edge_length);
    std::fprintf(f, "dt %.17g\nT %.17g\nbuckling %d\n", mdremote.timestep, T, o.B == 'y' ? 1 : 0);
    std::fprintf(f, "total_area %.17g\ntotal_volume %.17g\n", boundary.total_area, boundary.total_volume);
    std::fprintf(f, "netMeshKE %.17g\npenergy %.17g\nenergy %.17g\n", boundary.netMeshKE, boundary.penergy,
                 boundary.energy);

    // Energy components: compute_energy keeps them in locals (src/interface.cpp:990-1033), so
    // re-evaluate each with the reference's own member/pair functions and check the sum against
    // boundary.penergy, which compute_energy did store.
    long double bE = 0, sE = 0, ljE = 0, esE = 0;
    for (size_t i = 0; i < boundary.E.size(); i++)
        bE += boundary.bkappa *
              (1 - boundary.E[i].itsS / (4 * boundary.E[i].itsF[0]->itsarea * boundary.E[i].itsF[1]->itsarea));
    if (o.B == 'n') {
        for (size_t i = 0; i < boundary.E.size(); i++) {
            double stretched = boundary.E[i].length() - boundary.E[i].l0;
            sE += 0.5 * boundary.sconstant * stretched * stretched / (boundary.E[i].l0 * boundary.E[i].l0);
        }
        sE = sE * boundary.avg_edge_length * boundary.avg_edge_length;
    } else {
        for (size_t i = 0; i < boundary.E.size(); i++) {
            double stretched = boundary.E[i].length() - boundary.avg_edge_length;
            sE += 0.5 * boundary.sconstant * stretched * stretched;
        }
    }
    double tE = boundary.sigma_a * boundary.total_area;
    double volE = boundary.sigma_v * ((boundary.total_volume - boundary.ref_volume) *
                                      (boundary.total_volume - boundary.ref_volume));
    for (size_t i = 0; i < N; i++) {
        long double lj_i = 0, es_i = 0;
        for (size_t j = 0; j < N; j++) {
            if (i == j) continue;
            lj_i += energy_lj_pair(boundary.V[i].posvec, boundary.V[j].posvec, boundary.lj_length_mesh_mesh,
                                   boundary.elj);
            es_i += energy_es_pair(boundary.V[i], boundary.V[j], boundary.em, boundary.inv_kappa_out, scalefactor);
        }
        ljE += 0.5L * lj_i;
        esE += 0.5L * es_i;
        for (size_t j = 0; j < counterions.size(); j++) {  // mesh-ion, counted once (src/interface.cpp:1059-1062)
            ljE += energy_lj_pair(boundary.V[i].posvec, counterions[j].posvec, boundary.lj_length_mesh_ions, boundary.elj);
            esE += energy_es_pair(boundary.V[i], counterions[j], boundary.em, boundary.inv_kappa_out, scalefactor);
        }
    }
    for (size_t i = 0; i < counterions.size(); i++) {  // ion-ion (src/interface.cpp:1073-1084)
        long double lj_i = 0, es_i = 0;
        for (size_t j = 0; j < counterions.size(); j++) {
            if (i == j) continue;
            lj_i += energy_lj_pair(counterions[i].posvec, counterions[j].posvec, counterions[0].diameter, boundary.elj);
            es_i += energy_es_pair(counterions[i], counterions[j], boundary.em, boundary.inv_kappa_out, scalefactor);
        }
        ljE += 0.5L * lj_i;
        esE += 0.5L * es_i;
    }
    long double psum = bE + sE + tE + volE + ljE + esE;
    std::fprintf(f, "bE %.21Lg\nsE %.21Lg\ntE %.17g\nvolE %.17g\nljE %.21Lg\nesE %.21Lg\npenergy_check %.21Lg\n", bE,
                 sE, tE, volE, ljE, esE, psum);
    bool wrote = (o.steps < 3) || (o.steps % mdremote.writedata == 0);  // src/newmd.cpp:174
    std::fprintf(f, "energy_fresh %d\n", wrote ? 1 : 0);
    if (wrote && fabsl(psum - (long double) boundary.penergy) > 1e-11L * (fabsl(psum) + 1)) {
        std::fprintf(stderr, "component sum %.17Lg != penergy %.17g\n", psum, boundary.penergy);
        return 3;
    }

    long double ke = compute_kinetic_energy(boundary.V);
    std::fprintf(f, "vertex_ke %.21Lg\n", ke);
    std::fprintf(f, "mesh_bath_ke %.17g\nmesh_bath_pe %.17g\nions_bath_ke %.17g\nions_bath_pe %.17g\n",
                 bath_kinetic_energy(mesh_bath), bath_potential_energy(mesh_bath), bath_kinetic_energy(ions_bath),
                 bath_potential_energy(ions_bath));
    std::fprintf(f, "chain %zu\n", mesh_bath.size());
    for (size_t j = 0; j < mesh_bath.size(); j++)
        std::fprintf(f, "mesh_bath %zu %.17g %.17g %d %.17g %.17g\n", j, mesh_bath[j].Q, mesh_bath[j].T,
                     mesh_bath[j].dof, mesh_bath[j].xi, mesh_bath[j].eta);
    for (size_t j = 0; j < ions_bath.size(); j++)
        std::fprintf(f, "ions_bath %zu %.17g %.17g %d %.17g %.17g\n", j, ions_bath[j].Q, ions_bath[j].T,
                     ions_bath[j].dof, ions_bath[j].xi, ions_bath[j].eta);

    for (size_t i = 0; i < N; i++) {
        VERTEX &V = boundary.V[i];
        std::fprintf(f, "q %zu %.17g\n", i, V.q);
        pv(f, "pos", i, V.posvec);
        pv(f, "vel", i, V.velvec);
        pv(f, "for", i, V.forvec);
        pv(f, "bfo", i, V.bforce);
        pv(f, "sfo", i, V.sforce);
        pv(f, "tfo", i, V.TForce);
        pv(f, "vfo", i, V.VolTForce);
        VECTOR3D pair = V.forvec - V.bforce - V.sforce - V.TForce - V.VolTForce;
        pv(f, "pfo", i, pair);
    }
    std::fprintf(f, "n_ions %zu\nion_diameter %.17g\nlj_length_mesh_ions %.17g\nnetIonKE %.17g\n", counterions.size(),
                 counterion_diameter, boundary.lj_length_mesh_ions, boundary.netIonKE);
    for (size_t i = 0; i < counterions.size(); i++) {
        PARTICLE &P = counterions[i];
        std::fprintf(f, "ionq %zu %.17g\n", i, P.q);
        pv(f, "ipos", i, P.posvec);
        pv(f, "ivel", i, P.velvec);
        pv(f, "ifor", i, P.forvec);
    }
    for (size_t i = 0; i < boundary.E.size(); i++)
        std::fprintf(f, "edge %zu %u %u %.17g %.17g\n", i, boundary.E[i].itsV[0]->index, boundary.E[i].itsV[1]->index,
                     boundary.E[i].l0, boundary.E[i].itsS);
    for (size_t i = 0; i < boundary.F.size(); i++)
        std::fprintf(f, "face %zu %u %u %u %.17g %.17g\n", i, boundary.F[i].itsV[0]->index,
                     boundary.F[i].itsV[1]->index, boundary.F[i].itsV[2]->index, boundary.F[i].itsarea,
                     boundary.F[i].subtends_volume);
    if (o.local) {
        // VERTEX::itsarea / itsnormal the way the movie writer refreshes them (src/newmd.cpp:232-235)
        for (size_t i = 0; i < boundary.F.size(); i++) boundary.F[i].compute_area_normal();
        for (size_t i = 0; i < N; i++) boundary.V[i].compute_area_normal();
        for (size_t i = 0; i < N; i++)
            std::fprintf(f, "varea %zu %.17g %.21Lg %.21Lg %.21Lg\n", i, boundary.V[i].itsarea, boundary.V[i].itsnormal.x,
                         boundary.V[i].itsnormal.y, boundary.V[i].itsnormal.z);
        // The reference's own local maps land in outfiles/local_*_E.off with 6 significant digits
        // (src/interface.cpp:1140-1337); the same sums are repeated here with the reference's pair / edge functions so
        // that they can be printed in full. oracle/gen_golden.py checks the two against each other.
        boundary.compute_local_energies(scalefactor);
        boundary.compute_local_energies_by_component();
        std::vector<double> es(boundary.F.size(), 0), el(boundary.F.size(), 0), be(boundary.F.size(), 0), st(boundary.F.size(), 0);
        for (size_t i = 0; i < N; i++)
            for (size_t j = i + 1; j < N; j++) {
                double cur_E = energy_es_pair(boundary.V[i], boundary.V[j], boundary.em, boundary.inv_kappa_out, scalefactor);
                for (size_t k = 0; k < boundary.V[i].itsF.size(); k++) es[boundary.V[i].itsF[k]->index] += cur_E;
                for (size_t k = 0; k < boundary.V[j].itsF.size(); k++) es[boundary.V[j].itsF[k]->index] += cur_E;
            }
        for (size_t i = 0; i < boundary.E.size(); i++) {
            EDGE &E = boundary.E[i];
            double stretched = (E.length() - E.l0);
            double senergy = 0.5 * boundary.sconstant * stretched * stretched / (E.l0 * E.l0);
            senergy *= boundary.avg_edge_length * boundary.avg_edge_length;
            double benergy = boundary.bkappa * (1 - E.itsS / (4 * E.itsF[0]->itsarea * E.itsF[1]->itsarea));
            for (size_t k = 0; k < E.itsF.size(); k++) {
                el[E.itsF[k]->index] += senergy;
                st[E.itsF[k]->index] += senergy;
            }
            for (size_t k = 0; k < E.itsF.size(); k++) {
                el[E.itsF[k]->index] += benergy;
                be[E.itsF[k]->index] += benergy;
            }
        }
        for (size_t i = 0; i < boundary.F.size(); i++)
            std::fprintf(f, "lface %zu %.17g %.17g %.17g %.17g\n", i, es[i], el[i], be[i], st[i]);
    }
    std::fclose(f);
    std::fprintf(stderr, "dumped step %d to %s (md wall %.3f s)\n", o.steps, o.out.c_str(), t1 - t0);
    return 0;
}
