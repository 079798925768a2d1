// Command-line driver over the host mirror: the part of the reference's main() that leads up to
// md_interface (src/main.cpp:196-294, 467), with the reference's option letters for the inputs that
// reach the path. Runs from a directory holding infiles/<3>_<D>.dat; series files go to outfiles/.
//
//   np_shape_lab_b200 -R 10 -q 600 -c 0.005 -t 1 -v 1 -b 40 -s 40 -S 25000 -D 4 -F n
#include <sys/stat.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "npsl_host.hpp"

int main(int argc, char **argv) {
    double R = 10, q = 600, conc = 0.005, t = 1, v = 1, b = 40, s = 40, p = 0.5, dt = 0.001, T = 0.001, Qm = 0.01, Qi = 0.01;
    int N = 1, steps = 25000, D = 4, C = 5, writedata = 500, anneal_freq = 50000, anneal_dura = 5000;  // src/main.cpp:148-168
    char B = 'n', Fflag = 'n', H = 'n', G = 'N', annealFlag = 'y';
    std::string mesh_file, outdir = "outfiles", externalPattern = "None", ions_file;
    int dump_q = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-h" || a == "--help") {
            std::puts("np_shape_lab_b200 [-R r] [-q q] [-c conc] [-t st] [-v vt] [-b bend] [-s stretch] [-B y|n] [-N 1|2] [-p frac]\n"
                      "                  [-H n|y|c] [-F n|y] [-E pattern] [-G N|V]\n"
                      "                  [-S steps] [-D disc] [-d dt] [-T temp] [-Q Qmesh] [-C chain] [-W writedata]\n"
                      "                  [-a y|n (anneal)] [-f anneal_freq] [-u anneal_dura]\n"
                      "                  [--mesh file.dat] [--out dir] [--ions counterions.xyz]\n"
                      "                  (needs a CUDA device; there is no CPU path)");
            return 0;
        }
        if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); return 2; }
        const char *val = argv[++i];
        if (a == "-R") R = atof(val); else if (a == "-q") q = atof(val); else if (a == "-c") conc = atof(val);
        else if (a == "-t") t = atof(val); else if (a == "-v") v = atof(val); else if (a == "-b") b = atof(val);
        else if (a == "-s") s = atof(val); else if (a == "-B") B = val[0]; else if (a == "-F") Fflag = val[0];
        else if (a == "-H") H = val[0]; else if (a == "-G") G = val[0]; else if (a == "-N") N = atoi(val);
        else if (a == "-p") p = atof(val); else if (a == "-S") steps = atoi(val); else if (a == "-D") D = atoi(val);
        else if (a == "-d") dt = atof(val); else if (a == "-T") T = atof(val); else if (a == "-Q") Qm = atof(val);
        else if (a == "-Y") Qi = atof(val); else if (a == "-C") C = atoi(val); else if (a == "-W") writedata = atoi(val);
        else if (a == "--mesh") mesh_file = val; else if (a == "--out") outdir = val;
        else if (a == "-E") externalPattern = val; else if (a == "--dump-q") dump_q = atoi(val);
        else if (a == "--ions") ions_file = val;
        else if (a == "-a") annealFlag = val[0]; else if (a == "-f") anneal_freq = atoi(val); else if (a == "-u") anneal_dura = atoi(val);
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    try {
        INTERFACE boundary;
        CONTROL mdremote;
        std::vector<PARTICLE> counterions;
        boundary.outdir = outdir;
        if (!outdir.empty()) mkdir(outdir.c_str(), 0755);
        // unit scalings, src/main.cpp:196-210
        boundary.sigma_a = (std::pow(10, -14) * (std::pow(R, 2) / unitenergy)) * t;
        boundary.sigma_v = ((1.01325 * std::pow(10, 5)) * std::pow(10, -27) * (std::pow(R, 6) / (unitenergy * std::pow(10, -7)))) * v;
        boundary.bkappa = b;
        boundary.sconstant = s;
        const double scalefactor = epsilon_water * lB_water / R;
        mdremote.steps = steps; mdremote.timestep = dt; mdremote.writedata = writedata;
        mdremote.anneal = annealFlag; mdremote.annealfreq = anneal_freq; mdremote.annealDuration = anneal_dura;
        mdremote.QAnnealFac = 1.00 + (mdremote.TAnnealFac / 10.0);
        boundary.ein = epsilon_water; boundary.eout = epsilon_water;
        boundary.set_up(R);
        const double z_out = 1;
        boundary.inv_kappa_out = conc == 0 ? 100000000 : (0.257 / std::sqrt(z_out * boundary.lB_out * R * conc)) / R;
        if (mesh_file.empty()) boundary.discretize(3, D); else boundary.discretize_file(mesh_file);
        if (externalPattern == "None") boundary.assign_random_q_values(q, 1.0, N, p, Fflag, H);  // src/main.cpp:244-247
        else boundary.assign_external_q_values(q, externalPattern);
        if (dump_q) {  // charges only (no device needed): one "%.17g" line per vertex
            for (size_t i = 0; i < boundary.V.size(); i++) std::printf("%.17g\n", boundary.V[i].q);
            return 0;
        }
        boundary.lj_length_mesh_mesh = 0.1 * boundary.avg_edge_length;
        const double counterion_diameter = 0.6 / R;                                                  // src/main.cpp:196
        boundary.lj_length_mesh_ions = ((2.0 * (0.5 * boundary.avg_edge_length) + counterion_diameter) / 2.0);  // :271-272
        if (!ions_file.empty()) read_counterions(ions_file, counterion_diameter, 1, counterions);
        boundary.elj = 1.0;
        boundary.dressup(0, 0);
        boundary.ref_area = boundary.total_area;
        boundary.ref_volume = boundary.total_volume;
        // thermostat chains, src/main.cpp:276-290
        std::vector<THERMOSTAT> mesh_bath, ions_bath;
        const double n3 = 3.0 * boundary.V.size();
        if (C == 1) {
            mesh_bath.push_back(THERMOSTAT(0, T, n3));
            ions_bath.push_back(THERMOSTAT(0, T, 3.0 * counterions.size()));
        } else {
            mesh_bath.push_back(THERMOSTAT(Qm, T, n3));
            ions_bath.push_back(THERMOSTAT(Qi, T, 3.0 * counterions.size()));
            while ((int) mesh_bath.size() != C - 1) mesh_bath.push_back(THERMOSTAT(Qm, T, 1));
            while ((int) ions_bath.size() != C - 1) ions_bath.push_back(THERMOSTAT(Qi, T, 1));
            mesh_bath.push_back(THERMOSTAT(0, T, n3));
            ions_bath.push_back(THERMOSTAT(0, T, 3.0 * counterions.size()));
        }
        for (size_t i = 0; i < boundary.V.size(); i++) boundary.V[i].m = 1.0;
        std::cout << "Number of vertices " << boundary.V.size() << ", edges " << boundary.E.size() << ", faces "
                  << boundary.F.size() << "\nDebye length " << boundary.inv_kappa_out << "\nInitial area "
                  << boundary.ref_area << ", volume " << boundary.ref_volume << "\nTotal charge ";
        double qt = 0;
        for (size_t i = 0; i < boundary.V.size(); i++) qt += boundary.V[i].q;
        std::cout << qt << std::endl;
        const double box_halflength = 1.0;  // src/main.cpp:194
        output_directory() = outdir;
        if (!outdir.empty()) {  // initial frame of the movie ("time step -1"), src/main.cpp:417
            boundary.compute_vertex_area_normals();
            interface_movie(0, boundary.V, counterions, box_halflength);
        }
        const auto t0 = std::chrono::steady_clock::now();
        md_interface(boundary, counterions, mesh_bath, ions_bath, mdremote, G, B, 'L', scalefactor, box_halflength);
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        boundary.compute_local_energies(scalefactor);  // src/main.cpp:468-469
        boundary.compute_local_energies_by_component();
        const npsl_energies &e = boundary.parts;
        const double E = boundary.energy + e.mesh_bath_ke + e.mesh_bath_pe + e.ions_bath_ke + e.ions_bath_pe;
        std::string baths = "[";  // THERMOSTAT Q, T, dof, xi, eta of every link of the mesh chain (annealing rescales Q and T)
        for (size_t j = 0; j < mesh_bath.size(); j++) {
            char buf[160];
            std::snprintf(buf, sizeof buf, "%s[%.17g, %.17g, %.17g, %.17g, %.17g]", j ? ", " : "", mesh_bath[j].Q, mesh_bath[j].T,
                          (double) mesh_bath[j].dof, mesh_bath[j].xi, mesh_bath[j].eta);
            baths += buf;
        }
        baths += "]";
        std::printf("{\"vertices\": %zu, \"steps\": %d, \"wall_s\": %.6f, \"A\": %.17g, \"U\": %.17g, \"E\": %.17g, "
                    "\"KE_plus_PE\": %.17g, \"mesh_bath\": %s}\n",
                    boundary.V.size(), steps, wall, boundary.total_area, boundary.penergy, E, boundary.energy, baths.c_str());
    } catch (const npsl_error &err) {
        std::fprintf(stderr, "np_shape_lab_b200: %s (status %d)\n", err.what(), err.status);
        return 1;
    }
    return 0;
}
