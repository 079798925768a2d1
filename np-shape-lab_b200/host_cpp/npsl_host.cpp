// Bodies of the host mirror (npsl_host.hpp): set-up on the host exactly as main() does it before the
// time loop (src/main.cpp:196-294), the five path functions as thin adapters over the C ABI.
#include "npsl_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <utility>

using std::string;
using std::vector;

const double unitenergy = 1.3807 * std::pow(10.0, -16) * room_temperature;

namespace {

void check(npsl_ctx *ctx, int st, const char *what) {
    if (st != NPSL_OK) throw npsl_error(st, string(what) + ": " + npsl_last_error(ctx));
}


void append_line(const string &dir, const char *file, const string &line) {
    if (dir.empty()) return;
    std::ofstream out((dir + "/" + file).c_str(), std::ios::app);
    out << line << "\n";
}

template <class T>
string cell(const T &v) {  // one setw(15) column in the reference's default stream format
    std::ostringstream s;
    s << std::setw(15) << v;
    return s.str();
}

void fill_links(npsl_bath_link *dst, const vector<THERMOSTAT> &src) {
    for (size_t j = 0; j < src.size() && j < NPSL_MAX_CHAIN; j++) {
        dst[j].Q = src[j].Q; dst[j].T = src[j].T; dst[j].dof = src[j].dof; dst[j].xi = src[j].xi; dst[j].eta = src[j].eta;
    }
}

void read_links(vector<THERMOSTAT> &dst, const npsl_bath_link *src) {
    for (size_t j = 0; j < dst.size() && j < NPSL_MAX_CHAIN; j++) {
        dst[j].Q = src[j].Q; dst[j].T = src[j].T; dst[j].dof = src[j].dof; dst[j].xi = src[j].xi; dst[j].eta = src[j].eta;
    }
}

}  // namespace

double VECTOR3D::GetMagnitude() const { return std::sqrt(x * x + y * y + z * z); }
double EDGE::length() const { return (itsV[0]->posvec - itsV[1]->posvec).GetMagnitude(); }  // src/edge.cpp:24-31

INTERFACE::INTERFACE() { std::memset(&parts, 0, sizeof parts); }
INTERFACE::~INTERFACE() { drop_context(); }

void INTERFACE::drop_context() {
    if (ctx) npsl_destroy(ctx);
    ctx = nullptr;
}

void INTERFACE::set_up(double unit_radius_sphere) {
    em = 0.5 * (ein + eout);
    lB_out = (lB_water * epsilon_water / eout) / unit_radius_sphere;
}

void INTERFACE::discretize(unsigned int disc1, unsigned int disc2) {
    std::ostringstream name;
    name << "infiles/" << disc1 << "_" << disc2 << ".dat";
    discretize_file(name.str());
}

// The reference's text format: three sections, each opened by six header tokens
// ("# Number of Vertices= n vertices"); records "idx x y z q", "idx v1 v2 tag", "idx v1 v2 v3 sign tag".
void INTERFACE::discretize_file(const string &path) {
    std::ifstream in(path.c_str());
    if (!in) throw npsl_error(NPSL_ERR_ARG, "mesh file could not be opened: " + path);
    drop_context();
    V.clear(); E.clear(); F.clear();
    string tok;
    unsigned int n = 0, a, b, c, idx;
    double x, y, z, q0;
    in >> tok >> tok >> tok >> tok >> n >> tok;
    number_of_vertices = n;
    V.resize(n);
    for (unsigned int i = 0; i < n; i++) {
        in >> idx >> x >> y >> z >> q0;
        V.at(idx).posvec = VECTOR3D(x, y, z);
        V[idx].index = idx;
        V[idx].q = q0;
    }
    in >> tok >> tok >> tok >> tok >> n >> tok;
    number_of_edges = n;
    E.resize(n);
    std::map<std::pair<unsigned, unsigned>, EDGE *> by_ends;
    for (unsigned int i = 0; i < n; i++) {
        in >> idx >> a >> b >> tok;
        EDGE &e = E.at(idx);
        e.index = idx;
        e.itsV.push_back(&V.at(a));
        e.itsV.push_back(&V.at(b));
        V[a].itsE.push_back(&e);
        V[b].itsE.push_back(&e);
        by_ends[std::make_pair(std::min(a, b), std::max(a, b))] = &e;
    }
    in >> tok >> tok >> tok >> tok >> n >> tok;
    number_of_faces = n;
    F.resize(n);
    int sign;
    for (unsigned int i = 0; i < n; i++) {
        in >> idx >> a >> b >> c >> sign >> tok;
        FACE &f = F.at(idx);
        f.index = idx;
        const unsigned w[3] = {sign == 1 ? a : c, b, sign == 1 ? c : a};  // stored winding
        for (int k = 0; k < 3; k++) f.itsV.push_back(&V.at(w[k]));
        V[a].itsF.push_back(&f); V[b].itsF.push_back(&f); V[c].itsF.push_back(&f);
        const unsigned side[3][2] = {{w[1], w[2]}, {w[2], w[0]}, {w[0], w[1]}};  // edge opposite corner k
        for (int k = 0; k < 3; k++) {
            auto it = by_ends.find(std::make_pair(std::min(side[k][0], side[k][1]), std::max(side[k][0], side[k][1])));
            if (it == by_ends.end()) throw npsl_error(NPSL_ERR_ARG, "face side without an edge record");
            f.itsE.push_back(it->second);
            it->second->itsF.push_back(&f);
        }
    }
    if (!in) throw npsl_error(NPSL_ERR_ARG, "mesh file truncated: " + path);
    avg_edge_length = 0;
    for (size_t i = 0; i < E.size(); i++) {
        E[i].l0 = E[i].length();
        avg_edge_length += E[i].l0;
    }
    avg_edge_length /= E.size();
}

namespace {
// vertex areas (a third of the incident faces) and the total area at the current positions
double vertex_areas(const INTERFACE &b, vector<double> &varea) {
    vector<double> farea(b.F.size());
    double total = 0;
    for (size_t f = 0; f < b.F.size(); f++) {
        const VECTOR3D d = (b.F[f].itsV[1]->posvec - b.F[f].itsV[0]->posvec) % (b.F[f].itsV[2]->posvec - b.F[f].itsV[1]->posvec);
        farea[f] = 0.5 * d.GetMagnitude();
        total += farea[f];
    }
    varea.assign(b.V.size(), 0.0);
    for (size_t f = 0; f < b.F.size(); f++)
        for (int k = 0; k < 3; k++) varea[b.F[f].itsV[k]->index] += farea[f];
    for (size_t i = 0; i < varea.size(); i++) varea[i] /= 3.0;
    return total;
}

// "Scale the vertices' charges to achieve the target net charge exactly", src/interface.cpp:776-798, 845-861
void rescale_charges(vector<VERTEX> &V, double q_target) {
    double actual = 0;
    for (size_t i = 0; i < V.size(); i++) actual += V[i].q;
    for (size_t i = 0; i < V.size(); i++) V[i].q = V[i].q * (q_target / actual);
}
}  // namespace

// INTERFACE::assign_random_q_values, src/interface.cpp:339-798. Vertices are ranked by (z, index); the vertex of rank i
// receives the area share of list entry i (vertex i unless shuffled); -N 1 uniform or cube (-H c), -N 2 Janus or
// yin-yang (-H y), -N 3..17 stripes about the z axis with the even bands charged; -F y shuffles the occupancy and area
// lists with srand(123456) + std::random_shuffle exactly like the reference (same libstdc++ / glibc stream); the total
// is then rescaled to an integer number of charges.
void INTERFACE::assign_random_q_values(double q_strength, double alpha, int num_divisions, double frac_patch,
                                       char randomFlag, char functionFlag) {
    if (num_divisions < 1 || num_divisions > 17) throw npsl_error(NPSL_ERR_ARG, "num_divisions must be in [1, 17]");
    const unsigned int n = (unsigned int) V.size();
    vector<double> areaList;
    const double total = vertex_areas(*this, areaList);
    vector<std::pair<double, int>> perm(n);
    for (unsigned int i = 0; i < n; i++) perm[i] = std::make_pair(V[i].posvec.z, (int) i);
    std::sort(perm.begin(), perm.end());
    vector<int> state;
    for (unsigned int i = 1; i <= alpha * n; i++) state.push_back(1);
    for (unsigned int i = (unsigned int) state.size(); i < n; i++) state.push_back(0);
    if (randomFlag == 'y') {
        srand(123456);
        std::random_shuffle(state.begin(), state.end());
        std::random_shuffle(areaList.begin(), areaList.end());
    }
    auto share = [&](unsigned int i) { return q_strength * areaList[i] / total; };
    auto at = [&](unsigned int i) -> VERTEX & { return V[perm[i].second]; };
    const unsigned int nVertPerPatch = (unsigned int) (frac_patch * V.size());
    const int N = num_divisions;
    if (N == 1) {
        for (unsigned int i = 0; i < n; i++) {
            const VECTOR3D &p = at(i).posvec;
            const bool on = functionFlag == 'c'
                                ? (p.x * p.x + p.y * p.y <= 0.25) || (p.y * p.y + p.z * p.z <= 0.25) || (p.x * p.x + p.z * p.z <= 0.25)
                                : state[i] == 1;
            at(i).q = on ? share(i) : 0;
        }
    } else if (N == 2) {
        unsigned int i;
        for (i = 0; i < nVertPerPatch; i++) at(i).q = share(i);
        for (; i < n; i++) at(i).q = 0;
        if (functionFlag == 'y')
            for (i = 0; i < n; i++) {
                const VECTOR3D &p = at(i).posvec;
                if (p.z <= 0 && (p.x - 0.5) * (p.x - 0.5) + p.z * p.z <= 0.25) at(i).q = 0;
                if (p.z >= 0 && (p.x + 0.5) * (p.x + 0.5) + p.z * p.z <= 0.25) at(i).q = share(i);
            }
    } else {
        unsigned int i = 0;
        for (int k = 1; k < N; k++) {  // the reference's bounds "number_of_vertices * k*(alpha / num_divisions)"
            const double bound = k == 1 ? n * (alpha / N) : (n * (unsigned int) k) * (alpha / N);
            for (; i < bound && i < n; i++) at(i).q = (k % 2 == 1) ? share(i) : 0;
        }
        for (; i < n; i++) at(i).q = (N % 2 == 1) ? share(i) : 0;
    }
    if (q_strength == 0) {
        for (unsigned int i = 0; i < n; i++) V[i].q = 0;
        return;
    }
    double q_target;
    if (N == 1) q_target = std::round(q_strength);
    else if (N == 2) q_target = std::round(frac_patch * q_strength);
    else q_target = std::round(std::ceil(0.5 * N) * frac_patch * q_strength);
    rescale_charges(V, q_target);
}

// INTERFACE::assign_external_q_values, src/interface.cpp:818-862: "index value" rows of
// infiles/Charge_Patterns/<externalPattern>.dat; vertices the file does not list keep the charge column of the mesh file.
void INTERFACE::assign_external_q_values(double q_strength, std::string externalPattern) {
    std::ifstream in(("infiles/Charge_Patterns/" + externalPattern + ".dat").c_str());
    if (!in) {
        std::cout << "Charge assignment file could not be opened.  Mesh will be uncharged." << std::endl;
        for (size_t i = 0; i < V.size(); i++) V[i].q = 0;
    } else {
        std::cout << "Charge assignment file opened successfully.  Charging the mesh." << std::endl;
        vector<double> areaList;
        const double total = vertex_areas(*this, areaList);
        int col1;
        double col2;
        while (in >> col1 >> col2) V.at(col1).q = (q_strength / total) * col2;
    }
    if (q_strength == 0) {
        for (size_t i = 0; i < V.size(); i++) V[i].q = 0;
        return;
    }
    double actual = 0;
    for (size_t i = 0; i < V.size(); i++) actual += V[i].q;
    if (actual != 0) rescale_charges(V, std::round(actual));
}

npsl_ctx *INTERFACE::context(double scalefactor, char bucklingFlag, const vector<THERMOSTAT> *mesh_bath,
                             const vector<THERMOSTAT> *ions_bath, double dt) {
    const double key[10] = {scalefactor, (double) bucklingFlag, bkappa, sconstant, sigma_a, sigma_v, ref_volume,
                            inv_kappa_out, dt, (double) (mesh_bath ? mesh_bath->size() : 0)};
    // calls that carry no thermostat (force_calculation, compute_energy, dressup) reuse whatever context the
    // last md_interface / force call created for the same physics
    if (ctx && std::memcmp(key, ctx_key, (mesh_bath ? 10 : 8) * sizeof(double)) == 0) return ctx;
    drop_context();
    vector<int32_t> ev(2 * E.size()), fv(3 * F.size());
    vector<double> l0(E.size());
    for (size_t e = 0; e < E.size(); e++) {
        ev[2 * e] = E[e].itsV[0]->index; ev[2 * e + 1] = E[e].itsV[1]->index;
        l0[e] = E[e].l0;
    }
    for (size_t f = 0; f < F.size(); f++)
        for (int k = 0; k < 3; k++) fv[3 * f + k] = F[f].itsV[k]->index;
    npsl_mesh_desc md = {(int32_t) V.size(), (int32_t) E.size(), (int32_t) F.size(), ev.data(), fv.data(), l0.data()};
    npsl_params p;
    std::memset(&p, 0, sizeof p);
    p.scalefactor = scalefactor; p.em = em; p.inv_kappa_out = inv_kappa_out; p.elj = elj;
    p.lj_length_mesh_mesh = lj_length_mesh_mesh; p.bkappa = bkappa; p.sconstant = sconstant;
    p.sigma_a = sigma_a; p.sigma_v = sigma_v; p.ref_volume = ref_volume; p.avg_edge_length = avg_edge_length;
    p.dt = dt; p.buckling = bucklingFlag == 'y' ? 1 : 0;
    p.chain = mesh_bath ? (int32_t) mesh_bath->size() : 1;
    if (mesh_bath) fill_links(p.mesh_bath, *mesh_bath);
    if (ions_bath) fill_links(p.ions_bath, *ions_bath);
    check(nullptr, npsl_create(&md, &p, &ctx), "npsl_create");
    std::memcpy(ctx_key, key, sizeof key);
    push_state();
    return ctx;
}

void INTERFACE::push_ions(vector<PARTICLE> &ions) {
    const size_t m = ions.size();
    vector<double> p(3 * m), v(3 * m), q(m);
    for (size_t i = 0; i < m; i++) {
        p[3 * i] = ions[i].posvec.x; p[3 * i + 1] = ions[i].posvec.y; p[3 * i + 2] = ions[i].posvec.z;
        v[3 * i] = ions[i].velvec.x; v[3 * i + 1] = ions[i].velvec.y; v[3 * i + 2] = ions[i].velvec.z;
        q[i] = ions[i].q;
    }
    check(ctx, npsl_set_ions(ctx, (int) m, p.data(), v.data(), q.data(), m ? ions[0].diameter : 0.0, lj_length_mesh_ions),
          "npsl_set_ions");
}

void INTERFACE::pull_ions(vector<PARTICLE> &ions) {
    const size_t m = ions.size();
    if (!m) return;
    vector<double> p(3 * m), v(3 * m), f(3 * m);
    check(ctx, npsl_download_ions(ctx, p.data(), v.data(), f.data()), "npsl_download_ions");
    for (size_t i = 0; i < m; i++) {
        ions[i].posvec = VECTOR3D(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        ions[i].velvec = VECTOR3D(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        ions[i].forvec = VECTOR3D(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
    }
}

void INTERFACE::push_state() {
    const size_t n = V.size();
    buf_a.resize(3 * n); buf_b.resize(3 * n); buf_c.resize(n);
    for (size_t i = 0; i < n; i++) {
        buf_a[3 * i] = V[i].posvec.x; buf_a[3 * i + 1] = V[i].posvec.y; buf_a[3 * i + 2] = V[i].posvec.z;
        buf_b[3 * i] = V[i].velvec.x; buf_b[3 * i + 1] = V[i].velvec.y; buf_b[3 * i + 2] = V[i].velvec.z;
        buf_c[i] = V[i].q;
    }
    check(ctx, npsl_upload_state(ctx, buf_a.data(), buf_b.data(), buf_c.data()), "npsl_upload_state");
}

void INTERFACE::pull_state() {
    const size_t n = V.size();
    vector<double> f(3 * n);
    buf_a.resize(3 * n); buf_b.resize(3 * n);
    check(ctx, npsl_download_state(ctx, buf_a.data(), buf_b.data(), f.data()), "npsl_download_state");
    for (size_t i = 0; i < n; i++) {
        V[i].posvec = VECTOR3D(buf_a[3 * i], buf_a[3 * i + 1], buf_a[3 * i + 2]);
        V[i].velvec = VECTOR3D(buf_b[3 * i], buf_b[3 * i + 1], buf_b[3 * i + 2]);
        V[i].forvec = VECTOR3D(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
    }
}

void INTERFACE::pull_forces() {
    const size_t n = V.size();
    vector<double> f(3 * n), c[5];
    check(ctx, npsl_download_state(ctx, nullptr, nullptr, f.data()), "npsl_download_state");
    for (int k = 0; k < 5; k++) c[k].resize(3 * n);
    check(ctx, npsl_download_force_components(ctx, c[0].data(), c[1].data(), c[2].data(), c[3].data(), c[4].data()),
          "npsl_download_force_components");
    for (size_t i = 0; i < n; i++) {
        V[i].forvec = VECTOR3D(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
        V[i].bforce = VECTOR3D(c[1][3 * i], c[1][3 * i + 1], c[1][3 * i + 2]);
        V[i].sforce = VECTOR3D(c[2][3 * i], c[2][3 * i + 1], c[2][3 * i + 2]);
        V[i].TForce = VECTOR3D(c[3][3 * i], c[3][3 * i + 1], c[3][3 * i + 2]);
        V[i].VolTForce = VECTOR3D(c[4][3 * i], c[4][3 * i + 1], c[4][3 * i + 2]);
    }
    vector<double> fa(F.size()), fvv(F.size()), fd(3 * F.size()), es(E.size());
    check(ctx, npsl_download_mesh_fields(ctx, fa.data(), fvv.data(), fd.data(), es.data()), "npsl_download_mesh_fields");
    for (size_t k = 0; k < F.size(); k++) {
        F[k].itsarea = fa[k]; F[k].subtends_volume = fvv[k];
        F[k].itsdirection = VECTOR3D(fd[3 * k], fd[3 * k + 1], fd[3 * k + 2]);
    }
    for (size_t k = 0; k < E.size(); k++) E[k].itsS = es[k];
}

void INTERFACE::dressup(double la, double lv) {
    lambda_a = la; lambda_v = lv;
    if (!ctx) {
        // before the first force call (src/main.cpp:262,421): a context with neutral physics suffices for geometry
        inv_kappa_out = inv_kappa_out > 0 ? inv_kappa_out : 1.0;
        context(epsilon_water * lB_water, 'n');
    } else {
        push_state();
    }
    check(ctx, npsl_dressup(ctx, &total_area, &total_volume), "npsl_dressup");
    vector<double> fa(F.size()), fvv(F.size()), fd(3 * F.size());
    check(ctx, npsl_download_mesh_fields(ctx, fa.data(), fvv.data(), fd.data(), nullptr), "npsl_download_mesh_fields");
    for (size_t k = 0; k < F.size(); k++) {
        F[k].itsarea = fa[k]; F[k].subtends_volume = fvv[k];
        F[k].itsdirection = VECTOR3D(fd[3 * k], fd[3 * k + 1], fd[3 * k + 2]);
    }
    for (size_t i = 0; i < V.size(); i++) {  // VERTEX::itsarea = a third of the incident faces
        double a = 0;
        for (size_t k = 0; k < V[i].itsF.size(); k++) a += V[i].itsF[k]->itsarea;
        V[i].itsarea = a / 3.0;
    }
}

void INTERFACE::compute_energy(int num, vector<PARTICLE> &counterions, const double scalefactor, char bucklingFlag) {
    (void) counterions;  // on the device since the last force call / md_interface
    context(scalefactor, bucklingFlag);
    check(ctx, npsl_energy(ctx, &parts), "npsl_energy");
    netMeshKE = parts.kinetic; netIonKE = parts.ions_kinetic; penergy = parts.potential; energy = parts.energy;
    total_area = parts.total_area; total_volume = parts.total_volume;
    std::ostringstream row;  // "num KE bE sE tE ljE esE volE", src/interface.cpp:987-1113
    row << num << " " << parts.kinetic << " " << parts.bending << " " << parts.stretching << " " << parts.tension << " "
        << parts.lj << " " << parts.electrostatic << " " << parts.volume;
    append_line(outdir, "energy_in_parts_kE_bE_sE_tE_ljE_esE.dat", row.str());
}

void INTERFACE::compute_vertex_area_normals() {
    if (!ctx) throw npsl_error(NPSL_ERR_STATE, "compute_vertex_area_normals before the first dressup / force call");
    vector<double> a(V.size()), n(3 * V.size());
    check(ctx, npsl_vertex_geometry(ctx, a.data(), n.data()), "npsl_vertex_geometry");
    for (size_t i = 0; i < V.size(); i++) {
        V[i].itsarea = a[i];
        V[i].itsnormal = VECTOR3D(n[3 * i], n[3 * i + 1], n[3 * i + 2]);
    }
}

namespace {
// One local_*_E.off file: header, vertex positions, then per face "3 i j k r g b a" with colormap(E) = (0, E, 1 - E, 1)
// (src/interface.cpp:12-18, 1141-1190), everything in the stream's default format like the reference.
void write_local_off(const INTERFACE &b, const string &file, const char *label, const vector<double> &e) {
    double lo = 1e100, hi = -1e100;
    for (size_t i = 0; i < e.size(); i++) { lo = std::min(lo, e[i]); hi = std::max(hi, e[i]); }
    std::cout << label << " range" << "\t" << lo << "\t" << hi << std::endl;
    if (b.outdir.empty()) return;
    std::ofstream out((b.outdir + "/" + file).c_str(), std::ios::out);
    out << "OFF\n" << b.V.size() << " " << b.F.size() << " " << b.E.size() << "\n";
    for (size_t i = 0; i < b.V.size(); i++) out << b.V[i].posvec.x << " " << b.V[i].posvec.y << " " << b.V[i].posvec.z << "\n";
    for (size_t i = 0; i < b.F.size(); i++)
        out << "3 " << b.F[i].itsV[0]->index << " " << b.F[i].itsV[1]->index << " " << b.F[i].itsV[2]->index << " " << 0.0 << " "
            << e[i] << " " << 1 - e[i] << " " << 1.0 << "\n";
}
}  // namespace

void INTERFACE::compute_local_energies(const double scalefactor) {
    if (!ctx) throw npsl_error(NPSL_ERR_STATE, "compute_local_energies before the first force call");
    (void) scalefactor;  // part of the context since the first force call
    vector<double> es(F.size()), el(F.size());
    check(ctx, npsl_local_energies(ctx, es.data(), el.data(), nullptr, nullptr), "npsl_local_energies");
    write_local_off(*this, "local_electrostatic_E.off", "electrostatic", es);
    write_local_off(*this, "local_elastic_E.off", "elastic", el);
}

void INTERFACE::compute_local_energies_by_component() {
    if (!ctx) throw npsl_error(NPSL_ERR_STATE, "compute_local_energies_by_component before the first force call");
    vector<double> be(F.size()), st(F.size());
    check(ctx, npsl_local_energies(ctx, nullptr, nullptr, be.data(), st.data()), "npsl_local_energies");
    write_local_off(*this, "local_stretching_E.off", "stretching", st);
    write_local_off(*this, "local_bending_E.off", "bending", be);
}

string &output_directory() {
    static string dir = "outfiles";
    return dir;
}

void interface_movie(int num, vector<VERTEX> &V, vector<PARTICLE> &counterions, double box_halflength) {
    if (output_directory().empty()) return;
    std::ofstream outdump((output_directory() + "/p.lammpstrj").c_str(), std::ios::app);
    outdump << std::setprecision(12);
    outdump << std::fixed;
    outdump << "ITEM: TIMESTEP" << std::endl;
    outdump << num - 1 << std::endl;
    outdump << "ITEM: NUMBER OF ATOMS" << std::endl;
    outdump << V.size() + counterions.size() << std::endl;
    outdump << "ITEM: BOX BOUNDS" << std::endl;
    for (int k = 0; k < 3; k++) outdump << -box_halflength << "\t" << box_halflength << std::endl;
    outdump << "ITEM: ATOMS index type x y z Vq Va Vq/a" << std::endl;
    for (size_t i = 0; i < V.size(); i++)
        outdump << i << "\t" << (V[i].q > 0 ? "1" : "-1") << "\t" << V[i].posvec.x << "\t" << V[i].posvec.y << "\t" << V[i].posvec.z
                << "\t" << V[i].q << "\t" << V[i].itsarea << "\t" << (V[i].q / V[i].itsarea) << std::endl;
    for (size_t i = 0; i < counterions.size(); i++)
        outdump << i + V.size() << "\t" << 2 << "\t" << counterions[i].posvec.x << "\t" << counterions[i].posvec.y << "\t"
                << counterions[i].posvec.z << "\t" << counterions[i].q << "\t" << 0 << "\t" << 0 << std::endl;
}

void interface_off(int num, INTERFACE &dsphere) {
    if (output_directory().empty()) return;
    std::ostringstream ss;
    ss << output_directory() << "/" << num << ".off";
    std::ofstream outdump(ss.str().c_str(), std::ios::out);
    outdump << "OFF\n" << dsphere.number_of_vertices << " " << dsphere.number_of_faces << " " << dsphere.number_of_edges << "\n";
    for (size_t i = 0; i < dsphere.V.size(); i++)
        outdump << dsphere.V[i].posvec.x << " " << dsphere.V[i].posvec.y << " " << dsphere.V[i].posvec.z << "\n";
    for (size_t i = 0; i < dsphere.F.size(); i++) {
        FACE &f = dsphere.F[i];
        outdump << "3 " << f.itsV[0]->index << " " << f.itsV[1]->index << " " << f.itsV[2]->index << " ";
        if (f.itsV[0]->q > 0 && f.itsV[1]->q > 0 && f.itsV[2]->q > 0) outdump << "255 0 0 1\n";
        else if (f.itsV[2]->q < 0 && f.itsV[1]->q < 0 && f.itsV[0]->q < 0) outdump << "0 0 255 1\n";
        else outdump << "0 0 0 0\n";
    }
}

void force_calculation_init(INTERFACE &boundary, vector<PARTICLE> &counterions, const double scalefactor, char bucklingFlag) {
    npsl_ctx *c = boundary.context(scalefactor, bucklingFlag);
    boundary.push_state();
    boundary.push_ions(counterions);
    check(c, npsl_force_init(c), "npsl_force_init");
    boundary.pull_forces();
    boundary.pull_ions(counterions);
}

void force_calculation(INTERFACE &boundary, vector<PARTICLE> &counterions, const double scalefactor, char bucklingFlag) {
    npsl_ctx *c = boundary.context(scalefactor, bucklingFlag);
    boundary.push_state();
    boundary.push_ions(counterions);
    check(c, npsl_force(c), "npsl_force");
    boundary.pull_forces();
    boundary.pull_ions(counterions);
}

long double compute_kinetic_energy(vector<PARTICLE> &counterions) {
    long double ke = 0;
    for (size_t i = 0; i < counterions.size(); i++) {
        counterions[i].ke = 0.5 * counterions[i].m * (counterions[i].velvec * counterions[i].velvec);
        ke += counterions[i].ke;
    }
    return ke;
}

void read_counterions(const string &path, double diameter, int valency, vector<PARTICLE> &counterions) {
    std::ifstream in(path.c_str());
    if (!in) throw npsl_error(NPSL_ERR_ARG, "counterion file could not be opened: " + path);
    size_t n = 0;
    string tag;
    in >> n >> tag;  // count, "counterions"
    counterions.clear();
    for (size_t i = 0; i < n; i++) {
        double x, y, z, r;
        in >> tag >> x >> y >> z >> r;
        if (!in) throw npsl_error(NPSL_ERR_ARG, "counterion file truncated: " + path);
        counterions.push_back(PARTICLE((int) i + 1, diameter, valency, valency * -1.0, 1.0, VECTOR3D(x, y, z)));
    }
}

long double compute_kinetic_energy(vector<VERTEX> &V) {
    long double ke = 0;
    for (size_t i = 0; i < V.size(); i++) {
        V[i].ke = 0.5 * V[i].m * (V[i].velvec * V[i].velvec);
        ke += V[i].ke;
    }
    return ke;
}

double bath_kinetic_energy(vector<THERMOSTAT> &bath) {
    double s = 0;
    for (size_t j = 0; j < bath.size(); j++) s += bath[j].kinetic_energy();
    return s;
}

double bath_potential_energy(vector<THERMOSTAT> &bath) {
    double s = 0;
    for (size_t j = 0; j < bath.size(); j++) s += bath[j].potential_energy();
    return s;
}

void initialize_velocities_to_zero(vector<VERTEX> &V, vector<PARTICLE> &ions) {
    for (size_t i = 0; i < V.size(); i++) V[i].velvec = VECTOR3D(0, 0, 0);
    for (size_t i = 0; i < ions.size(); i++) ions[i].velvec = VECTOR3D(0, 0, 0);
}

// The loop body runs on the device (one CUDA-graph launch per step); the host touches the state only
// when the reference writes (num < 3 or num % writedata == 0), anneals, or at the end.
void md_interface(INTERFACE &boundary, vector<PARTICLE> &counterions, vector<THERMOSTAT> &mesh_bath,
                  vector<THERMOSTAT> &ions_bath, CONTROL &mdremote, char geomConstraint, char bucklingFlag,
                  char constraintForm, const double scalefactor, double box_radius) {
    if (geomConstraint == 'V' && constraintForm == 'Q')
        throw npsl_error(NPSL_ERR_ARG, "the quadratic constraint form is not on the device path");
    if (geomConstraint != 'N' && geomConstraint != 'V')
        throw npsl_error(NPSL_ERR_ARG, "only the rigid volume constraint (-G V, linear form) is on the device path");
    if (mesh_bath.size() != ions_bath.size() || mesh_bath.empty() || mesh_bath.size() > NPSL_MAX_CHAIN)
        throw npsl_error(NPSL_ERR_ARG, "thermostat chains must have equal length in [1, NPSL_MAX_CHAIN]");
    initialize_velocities_to_zero(boundary.V, counterions);
    boundary.drop_context();
    npsl_ctx *c = boundary.context(scalefactor, bucklingFlag, &mesh_bath, &ions_bath, mdremote.timestep);
    if (geomConstraint == 'V') check(c, npsl_set_volume_constraint(c, 1), "npsl_set_volume_constraint");
    boundary.push_ions(counterions);
    check(c, npsl_force_init(c), "npsl_force_init");
    if (geomConstraint == 'V') check(c, npsl_constraint_init(c), "npsl_constraint_init");  // src/newmd.cpp:33-38
    npsl_bath_link mb[NPSL_MAX_CHAIN], ib[NPSL_MAX_CHAIN];
    const string &dir = boundary.outdir;

    auto sync_baths = [&]() {
        check(c, npsl_get_thermostat(c, mb, ib), "npsl_get_thermostat");
        read_links(mesh_bath, mb);
        read_links(ions_bath, ib);
    };
    double initNetEnergy = 0;
    int abortCounter = 0;
    auto write_row = [&](int num) -> double {
        boundary.compute_energy(num, counterions, scalefactor, bucklingFlag);
        sync_baths();
        const npsl_energies &e = boundary.parts;
        const double bke = e.mesh_bath_ke + e.ions_bath_ke, bpe = e.mesh_bath_pe + e.ions_bath_pe;
        const double global = boundary.energy + bke + bpe;
        std::ostringstream s;
        s << num << cell(boundary.netMeshKE + boundary.netIonKE) << cell(boundary.penergy) << cell(boundary.energy) << cell(global)
          << cell(bke) << cell(bpe);
        append_line(dir, "energy_nanomembrane.dat", s.str());
        std::ostringstream a, v, t;
        a << num << cell(boundary.total_area);
        v << num << cell(boundary.total_volume);
        t << num << cell(2 * boundary.netMeshKE / (mesh_bath[0].dof * kB));
        append_line(dir, "area.dat", a.str());
        append_line(dir, "volume.dat", v.str());
        append_line(dir, "temperature.dat", t.str());
        return global;
    };
    const double g0 = write_row(0);
    initNetEnergy = boundary.energy + boundary.parts.mesh_bath_ke + boundary.parts.mesh_bath_pe;  // src/newmd.cpp:62
    (void) g0;

    double fT = std::pow(1.0 / mdremote.TAnnealFac, 1.0 / mdremote.annealDuration);
    double fQ = std::pow(1.0 / mdremote.QAnnealFac, 1.0 / mdremote.annealDuration);
    auto anneal_step = [&](int n) {
        return mdremote.anneal == 'y' && n > mdremote.annealDuration &&
               (n % mdremote.annealfreq == 0 || n % mdremote.annealfreq < mdremote.annealDuration);
    };
    int num = 0;
    while (num < mdremote.steps) {
        // run on the device up to the next step at which the host has something to do
        int nxt = mdremote.steps;
        nxt = std::min(nxt, num < 2 ? num + 1 : (num / mdremote.writedata + 1) * mdremote.writedata);
        if (mdremote.moviefreq > 0) nxt = std::min(nxt, (num / mdremote.moviefreq + 1) * mdremote.moviefreq);
        if (mdremote.anneal == 'y') {
            if (anneal_step(num + 1)) nxt = std::min(nxt, num + 1);
            nxt = std::min(nxt, (num / mdremote.annealfreq + 1) * mdremote.annealfreq);
            if (num < mdremote.annealDuration + 1) nxt = std::min(nxt, mdremote.annealDuration + 1);
        }
        check(c, npsl_step(c, nxt - num), "npsl_step");
        num = nxt;
        if (num % mdremote.writedata == 0 || num < 3) {
            const double global = write_row(num);
            const double ratio = global / initNetEnergy;
            std::ostringstream d;
            d << num << "\t" << ratio << "\t" << global << "\t" << initNetEnergy << "\t" << abortCounter;
            append_line(dir, "net_Energy_Drift.dat", d.str());
            if (!(1.05 < ratio) && num < mdremote.annealfreq && ratio < 0.95) {  // src/newmd.cpp:205-220
                abortCounter++;
                if (abortCounter == 10)
                    throw npsl_error(NPSL_ERR_NUMERIC, "aborting: >5% net energy drift downwards prior to annealing");
            }
        }
        if (mdremote.moviefreq > 0 && num % mdremote.moviefreq == 0 && !dir.empty()) {  // snapshots, src/newmd.cpp:230-240
            boundary.pull_state();
            boundary.pull_ions(counterions);
            boundary.compute_vertex_area_normals();
            output_directory() = dir;
            if (mdremote.offfreq > 0 && num % mdremote.offfreq == 0) interface_off(num, boundary);
            interface_movie(num, boundary.V, counterions, box_radius);
        }
        if (anneal_step(num)) {  // gradual annealing, src/newmd.cpp:258-299
            sync_baths();
            if (mesh_bath[0].T >= 0.000000000000001) {
                if (num % mdremote.annealfreq == 0) {
                    if (num == mdremote.annealfreq) mdremote.TAnnealFac = 1.25;
                    else if (num <= 3 * mdremote.annealfreq) mdremote.TAnnealFac = 2;
                    else if (num <= 6 * mdremote.annealfreq) mdremote.TAnnealFac = 4;
                    else if (num <= 7 * mdremote.annealfreq) mdremote.TAnnealFac = 8;
                    else mdremote.TAnnealFac = 10;
                    mdremote.QAnnealFac = 1.0 + mdremote.TAnnealFac / 10.0;
                    fT = std::pow(1.0 / mdremote.TAnnealFac, 1.0 / mdremote.annealDuration);
                    fQ = std::pow(1.0 / mdremote.QAnnealFac, 1.0 / mdremote.annealDuration);
                }
                for (size_t b = 0; b < mesh_bath.size(); b++) {
                    mesh_bath[b].T = fT * mesh_bath[b].T;
                    mesh_bath[b].Q = fQ * mesh_bath[b].Q;
                }
                fill_links(mb, mesh_bath);
                fill_links(ib, ions_bath);
                check(c, npsl_set_thermostat(c, mb, ib), "npsl_set_thermostat");
            }
        }
    }
    boundary.pull_state();
    boundary.pull_ions(counterions);
    sync_baths();
}
