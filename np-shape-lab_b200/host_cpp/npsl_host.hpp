// C++ host mirror of the reference's call surface for the force / energy / integrator path.
//
// Same type and function names, argument meaning and side effects as the reference's in-process
// C++ surface; the bodies of the five path functions are replaced by calls over the C ABI
// (include/npsl.h -> libnpsl.so, hand-written sm_100a kernels):
//
//   force_calculation_init(INTERFACE&, vector<PARTICLE>&, scalefactor, bucklingFlag)   src/newforces.cpp:4
//   force_calculation(INTERFACE&, vector<PARTICLE>&, scalefactor, bucklingFlag)        src/newforces.cpp:137
//   INTERFACE::compute_energy(num, counterions, scalefactor, bucklingFlag)             src/interface.cpp:964
//   INTERFACE::dressup(lambda_a, lambda_v)                                             src/interface.cpp:33
//   md_interface(boundary, counterions, mesh_bath, ions_bath, mdremote, geomConstraint,
//                bucklingFlag, constraintForm, scalefactor, box_radius)                src/newmd.cpp:9
//
// Results are delivered the way the reference delivers them: by mutating the VERTEX / EDGE / FACE
// objects of `boundary`, its scalar members and the THERMOSTAT objects. The pointer-linked mesh
// (vector<EDGE*> itsE, vector<FACE*> itsF, ...) is kept for callers; the device works on the flat
// index tables built once from it. There is no CPU implementation behind these functions: without
// libnpsl.so and a CUDA device they throw npsl_error.
#ifndef NPSL_HOST_HPP
#define NPSL_HOST_HPP

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/npsl.h"

struct npsl_error : std::runtime_error {
    int status;
    npsl_error(int st, const std::string &what) : std::runtime_error(what), status(st) {}
};

// constants of src/utility.h:35-49
const double dcut2 = 1.25992104989487316476;
const double lB_water = 0.714;
const double epsilon_water = 80;
const double room_temperature = 298;
const double kB = 1;
extern const double unitenergy;  // 1.3807e-16 * room_temperature

// src/vector3d.h (the reference stores long double; the device path is FP64 throughout)
struct VECTOR3D {
    double x, y, z;
    VECTOR3D(double xx = 0, double yy = 0, double zz = 0) : x(xx), y(yy), z(zz) {}
    double GetMagnitude() const;
    VECTOR3D operator+(const VECTOR3D &v) const { return VECTOR3D(x + v.x, y + v.y, z + v.z); }
    VECTOR3D operator-(const VECTOR3D &v) const { return VECTOR3D(x - v.x, y - v.y, z - v.z); }
    VECTOR3D operator^(double s) const { return VECTOR3D(x * s, y * s, z * s); }                  // scale
    double operator*(const VECTOR3D &v) const { return x * v.x + y * v.y + z * v.z; }             // dot
    VECTOR3D operator%(const VECTOR3D &v) const {                                                 // cross
        return VECTOR3D(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x);
    }
};

class EDGE;
class FACE;
class INTERFACE;

// src/vertex.h:19-132 (the members the path reads or writes)
class VERTEX {
public:
    unsigned int index = 0;
    VECTOR3D posvec, velvec, forvec;
    VECTOR3D bforce, sforce, TForce, VolTForce;  // per-term forces, src/newforces.cpp:279-280
    VECTOR3D itsnormal;                          // area-weighted unit normal, src/vertex.cpp:7-18
    double q = 0, m = 1, ke = 0, itsarea = 0;
    std::vector<EDGE *> itsE;
    std::vector<FACE *> itsF;
    VERTEX(VECTOR3D p = VECTOR3D()) : posvec(p) {}
};

// src/edge.h:20-64
class EDGE {
public:
    unsigned int index = 0;
    std::vector<VERTEX *> itsV;
    std::vector<FACE *> itsF;
    double l0 = 0, itsS = 0;
    double length() const;
};

// src/face.h:16-60
class FACE {
public:
    unsigned int index = 0;
    std::vector<VERTEX *> itsV;
    std::vector<EDGE *> itsE;
    double itsarea = 0, subtends_volume = 0;
    VECTOR3D itsdirection;
};

// src/particle.h:13-91 (explicit counterions; all of one diameter, masses 1 like the reference's)
class PARTICLE {
public:
    int id = 0, valency = 0;
    VECTOR3D posvec, velvec, forvec;
    double q = 0, m = 1, diameter = 0, ke = 0;
    PARTICLE(int get_id = 0, double get_diameter = 0, int get_valency = 0, double get_charge = 0, double get_mass = 0,
             VECTOR3D get_position = VECTOR3D(0, 0, 0))
        : id(get_id), valency(get_valency), posvec(get_position), q(get_charge), m(get_mass), diameter(get_diameter) {}
};

// src/thermostat.h:8-71
class THERMOSTAT {
public:
    double Q, T;
    double dof;
    double xi, eta;
    THERMOSTAT(double Q_ = 0, double T_ = 0, double dof_ = 0, double xi_ = 0, double eta_ = 0, double = 0)
        : Q(Q_), T(T_), dof(dof_), xi(xi_), eta(eta_) {}
    double potential_energy() const { return dof * kB * T * eta; }
    double kinetic_energy() const { return 0.5 * Q * xi * xi; }
};

// src/control.h:6-42
class CONTROL {
public:
    double timestep = 0.001;
    int steps = 1000;
    char anneal = 'y';
    int annealfreq = 50000, annealDuration = 5000;
    double TAnnealFac = 1.25, QAnnealFac = 1.125;
    int moviestart = 1, offfreq = 2500, moviefreq = 1000, povfreq = 100000, writedata = 500;
};

// src/interface.h:20-150
class INTERFACE {
public:
    double ein = 0, eout = 0, em = 0, lB_out = 0, inv_kappa_out = 0;
    unsigned int number_of_vertices = 0, number_of_edges = 0, number_of_faces = 0;
    std::vector<VERTEX> V;
    std::vector<EDGE> E;
    std::vector<FACE> F;
    double ref_area = 0, ref_volume = 0;
    double total_area = 0, total_volume = 0;
    double avg_edge_length = 0;
    double lambda_a = 0, lambda_v = 0, lambda_l = 0;
    double bkappa = 0, sconstant = 0, sigma_a = 0, sigma_v = 0;
    double elj = 1, lj_length_mesh_mesh = 0, lj_length_mesh_ions = 0;
    double netMeshKE = 0, netIonKE = 0, penergy = 0, energy = 0;
    // components of the last compute_energy (the reference keeps them in locals and prints them)
    npsl_energies parts;
    std::string outdir = "outfiles";  // series files; empty string = write nothing

    INTERFACE();
    ~INTERFACE();
    INTERFACE(const INTERFACE &) = delete;
    INTERFACE &operator=(const INTERFACE &) = delete;

    void set_up(double unit_radius_sphere);                              // src/interface.cpp:20-29
    void discretize(unsigned int disc1, unsigned int disc2);             // src/interface.cpp:70-256 (infiles/<disc1>_<disc2>.dat)
    void discretize_file(const std::string &path);
    // charge patterns, src/interface.cpp:339-798: -N 1 uniform / cube (-H c), -N 2 Janus / yin-yang (-H y), -N 3..17 stripes,
    // fractional occupancy alpha, -F y shuffled; and patterns read from infiles/Charge_Patterns/<name>.dat (:818-862)
    void assign_random_q_values(double q_strength, double alpha, int num_divisions, double frac_patch, char randomFlag,
                                char functionFlag);
    void assign_external_q_values(double q_strength, std::string externalPattern);
    void dressup(double lambda_a, double lambda_v);                      // src/interface.cpp:33-67
    void compute_energy(int num, std::vector<PARTICLE> &counterions, const double scalefactor, char bucklingFlag);
    // local energy maps: <outdir>/local_electrostatic_E.off, local_elastic_E.off and local_bending_E.off,
    // local_stretching_E.off in the reference's format, per-face sums from the device (src/interface.cpp:1140-1337)
    void compute_local_energies(const double scalefactor);
    void compute_local_energies_by_component();
    // VERTEX::compute_area_normal for every vertex (src/vertex.cpp:7-18), as the movie writer refreshes them
    // (src/newmd.cpp:232-235): V[i].itsarea, V[i].itsnormal <- device
    void compute_vertex_area_normals();

    // ---- device plumbing (not in the reference) ----
    npsl_ctx *ctx = nullptr;
    npsl_ctx *context(double scalefactor, char bucklingFlag, const std::vector<THERMOSTAT> *mesh_bath = nullptr,
                      const std::vector<THERMOSTAT> *ions_bath = nullptr, double dt = 0.001);
    void drop_context();
    void push_ions(std::vector<PARTICLE> &ions);  // counterions -> device (positions, velocities, charges, LJ lengths)
    void pull_ions(std::vector<PARTICLE> &ions);  // PARTICLE::posvec / velvec / forvec <- device
    void push_state();              // posvec / velvec / q  -> device
    void pull_forces();             // forvec + per-term forces, EDGE::itsS, FACE fields <- device
    void pull_state();              // posvec / velvec / forvec <- device
private:
    double ctx_key[10] = {0};
    std::vector<double> buf_a, buf_b, buf_c;
};

void force_calculation_init(INTERFACE &boundary, std::vector<PARTICLE> &counterions, const double scalefactor,
                            char bucklingFlag);
void force_calculation(INTERFACE &boundary, std::vector<PARTICLE> &counterions, const double scalefactor,
                       char bucklingFlag);
void md_interface(INTERFACE &boundary, std::vector<PARTICLE> &counterions, std::vector<THERMOSTAT> &mesh_bath,
                  std::vector<THERMOSTAT> &ions_bath, CONTROL &mdremote, char geomConstraint, char bucklingFlag,
                  char constraintForm, const double scalefactor, double box_radius);

long double compute_kinetic_energy(std::vector<VERTEX> &V);              // src/newenergies.h:8-16 (host arrays)
long double compute_kinetic_energy(std::vector<PARTICLE> &counterions);  // src/newenergies.h:18-26
// counterions.xyz as INTERFACE::put_counterions writes it (src/interface.cpp:946-953: count, a comment line, then
// "C x y z |r|") read back into monovalent PARTICLEs of the given diameter -- the reference's own placement loop cannot
// terminate with its hard-wired box (src/main.cpp:194), so positions have to come from outside
void read_counterions(const std::string &path, double diameter, int valency, std::vector<PARTICLE> &counterions);
double bath_kinetic_energy(std::vector<THERMOSTAT> &bath);               // src/newenergies.h:28-36
double bath_potential_energy(std::vector<THERMOSTAT> &bath);             // src/newenergies.h:38-46
void initialize_velocities_to_zero(std::vector<VERTEX> &V, std::vector<PARTICLE> &ions);  // src/functions.cpp:40-47
// The reference's snapshot writers (host formatting only; areas come from compute_vertex_area_normals). They write
// into output_directory() ("outfiles" unless changed), like the reference's hard-wired "outfiles/".
void interface_movie(int num, std::vector<VERTEX> &V, std::vector<PARTICLE> &counterions, double box_halflength);  // src/functions.cpp:50-84  p.lammpstrj
void interface_off(int num, INTERFACE &dsphere);                                                                   // src/functions.cpp:134-167  <num>.off
std::string &output_directory();

#endif
