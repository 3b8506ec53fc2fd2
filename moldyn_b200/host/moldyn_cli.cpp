// moldyn_cli.cpp — driver with the flags of the reference's `moldyn_cli` (cli/src/args.rs:7-176), built on the
// C++ host mirror.  `solve` keeps the state resident on the GPU and only downloads at frame boundaries; with the
// default --frames-per-save 1 it writes a frame every step exactly like cli/src/commands.rs:182-190.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "moldyn.hpp"

using namespace moldyn;

namespace {

[[noreturn]] void usage(const std::string &msg)
{
    std::cerr << "error: " << msg << "\n"
              << "usage: moldyn_cli -f <dir> [--time] [--frames-per-save k] <command> [options]\n"
              << "commands: initialize, solve, solve-macro-parameters, check-impulse, particle-count,\n"
              << "          generate-default-potentials, set-potential, generate-velocities-histogram\n";
    std::exit(2);
}

struct Args {
    std::vector<std::string> v;
    size_t i = 0;
    bool done() const { return i >= v.size(); }
    const std::string &peek() const { return v[i]; }
    std::string next(const std::string &what)
    {
        if (done()) usage("missing value for " + what);
        return v[i++];
    }
    // clap `num_args = n, value_delimiter = ' '`: values may come as separate words or one quoted word
    std::vector<std::string> values(const std::string &what, size_t min_n, size_t max_n)
    {
        std::vector<std::string> out;
        while (!done() && out.size() < max_n) {
            const std::string &w = v[i];
            bool is_flag = w.size() > 1 && w[0] == '-' && !(std::isdigit((unsigned char)w[1]) || w[1] == '.');
            if (is_flag) break;
            size_t s = 0;
            while (s <= w.size()) {
                size_t e = w.find(' ', s);
                if (e == std::string::npos) e = w.size();
                if (e > s) out.push_back(w.substr(s, e - s));
                s = e + 1;
            }
            ++i;
        }
        if (out.size() < min_n) usage("not enough values for " + what);
        return out;
    }
};

bool is(const std::string &a, const char *s, const char *l) { return (s && a == s) || (l && a == l); }

size_t last_frame(const std::string &dir)  // cli/src/commands.rs:193-202
{
    DIR *d = opendir((dir + "/data").c_str());
    if (!d) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't read directory");
    long best = -1;
    while (dirent *e = readdir(d)) {
        std::string n = e->d_name;
        size_t dot = n.find('.');
        if (dot == std::string::npos || dot == 0) continue;
        best = std::max(best, std::atol(n.substr(0, dot).c_str()));
    }
    closedir(d);
    if (best < 0) throw Error(MD_ERR_INVALID_ARGUMENT, "no frames in data/");
    return (size_t)best;
}

// cli/src/commands.rs:43-80
void cmd_initialize(const std::string &dir, Args &a)
{
    std::string type, name;
    std::vector<std::string> size;
    double mass = 0, radius = 0, lattice = 0, temperature = 0;
    uint64_t seed = 0;
    while (!a.done()) {
        std::string o = a.next("option");
        if (is(o, "-t", "--crystal-cell-type")) type = a.next(o);
        else if (is(o, "-s", "--size")) size = a.values(o, 3, 3);
        else if (is(o, "-n", "--particle-name")) name = a.next(o);
        else if (is(o, "-m", "--particle-mass")) mass = std::stod(a.next(o));
        else if (is(o, "-r", "--particle-radius")) radius = std::stod(a.next(o));
        else if (is(o, "-l", "--lattice-cell")) lattice = std::stod(a.next(o));
        else if (is(o, "-T", "--temperature")) temperature = std::stod(a.next(o));
        else if (is(o, nullptr, "--seed")) seed = std::stoull(a.next(o));  // extension: reproducible velocities
        else usage("unknown option for initialize: " + o);
    }
    if (size.size() != 3 || name.empty() || !(mass > 0)) usage("initialize needs -t -s -n -m -r -l -T");
    initializer::UnitCell cell;
    if (type == "u") cell = initializer::UnitCell::U;
    else if (type == "fcc") cell = initializer::UnitCell::FCC;
    else usage("crystal cell type must be u or fcc");
    std::array<size_t, 3> g{std::stoul(size[0]), std::stoul(size[1]), std::stoul(size[2])};
    ParticleDatabase::add(0, name, mass, radius);
    size_t count = g[0] * g[1] * g[2] * (cell == initializer::UnitCell::FCC ? 4 : 1);
    Vector3 box{lattice * (double)g[0], lattice * (double)g[1], lattice * (double)g[2]};
    State st = initializer::initialize_particles({count}, box);
    initializer::initialize_particles_position(cell, st, 0, {0, 0, 0}, g, lattice);
    initializer::initialize_velocities_maxwell_boltzmann(st, temperature, 0, seed);
    StateToSave::from(st).save_to_file(dir, 0);
    ParticleDatabase::save_particles_data(dir);
}

// cli/src/commands.rs:82-191
void cmd_solve(const std::string &dir, Args &a, size_t frames_per_save)
{
    size_t state_number = 0, iteration_count = 0;
    bool have_s = false, have_c = false, have_t = false, use_potentials = false, exact = false;
    double dt = 0, temperature = 0, pressure = 0;
    bool have_T = false, have_P = false;
    std::string method, thermostat, barostat;
    std::vector<std::string> tparams, bparams;
    while (!a.done()) {
        std::string o = a.next("option");
        if (is(o, nullptr, "--threads-count")) a.next(o);  // CPU thread pool of the reference; no meaning here
        else if (is(o, "-s", "--state-number")) { state_number = std::stoul(a.next(o)); have_s = true; }
        else if (is(o, "-i", "--integrate-method")) method = a.next(o);
        else if (is(o, nullptr, "--custom-method")) a.next(o);
        else if (is(o, nullptr, "--barostat")) barostat = a.next(o);
        else if (is(o, nullptr, "--barostat-params")) bparams = a.values(o, 1, 4);
        else if (is(o, "-P", "--pressure")) { pressure = std::stod(a.next(o)); have_P = true; }
        else if (is(o, nullptr, "--thermostat")) thermostat = a.next(o);
        else if (is(o, nullptr, "--thermostat-params")) tparams = a.values(o, 1, 4);
        else if (is(o, "-T", "--temperature")) { temperature = std::stod(a.next(o)); have_T = true; }
        else if (is(o, "-p", "--use-potentials")) use_potentials = true;
        else if (is(o, "-c", "--iteration-count")) { iteration_count = std::stoul(a.next(o)); have_c = true; }
        else if (is(o, "-t", "--delta-time")) { dt = std::stod(a.next(o)); have_t = true; }
        else if (is(o, nullptr, "--exact")) exact = true;  // extension: MD_FORCE_EXACT
        else usage("unknown option for solve: " + o);
    }
    if (!have_s || !have_c || !have_t || method.empty()) usage("solve needs -s -i -c -t");
    if (method != "verlet-method") throw Error(MD_ERR_UNSUPPORTED, "Integrator::Custom is todo!() in the reference");

    StateToSave data = StateToSave::load_from_file(dir, state_number);
    ParticleDatabase::load_particles_data(dir);
    PotentialsDatabase db;
    State state = data.into_state();
    if (use_potentials) db.load_potentials_from_file(dir);

    Thermostat th{Thermostat::Berendsen};
    Barostat ba{Barostat::Berendsen};
    std::pair<Thermostat *, double> thp{&th, 0.0};
    std::pair<Barostat *, double> bap{&ba, 0.0};
    bool use_th = !thermostat.empty(), use_ba = !barostat.empty();
    if (use_th) {
        if (tparams.empty()) throw Error(MD_ERR_INVALID_ARGUMENT, "No thermostat parameters. Need tau");
        if (thermostat == "berendsen") th.kind = Thermostat::Berendsen;
        else if (thermostat == "nose-hoover") th.kind = Thermostat::NoseHoover;
        else throw Error(MD_ERR_UNSUPPORTED, "Thermostat::Custom is todo!() in the reference");
        th.tau = std::stod(tparams[0]);
        if (!have_T) throw Error(MD_ERR_INVALID_ARGUMENT, "No temperature was passed");
        thp.second = temperature;
    }
    if (use_ba) {
        if (bparams.size() < 2) throw Error(MD_ERR_INVALID_ARGUMENT, "No barostat parameters. Need beta and tau");
        if (barostat != "berendsen") throw Error(MD_ERR_UNSUPPORTED, "Barostat::Custom is todo!() in the reference");
        ba.beta = std::stod(bparams[0]);
        ba.tau = std::stod(bparams[1]);
        if (!have_P) throw Error(MD_ERR_INVALID_ARGUMENT, "No pressure was passed");
        bap.second = pressure;
    }

    Session s(0, exact);
    s.set_potentials(db, state);
    s.upload(state, false);
    s.update_force();  // commands.rs:102
    if (frames_per_save == 0) frames_per_save = 1;
    // frames go to a writer thread: the GPU steps the next chunk while the previous frame's CSV text is produced
    size_t frame = state_number, done = 0;
    FrameWriter writer(dir, 2);
    while (done < iteration_count) {
        s.download(state);
        writer.push(frame++, StateToSave::from(state));
        size_t chunk = std::min(frames_per_save, iteration_count - done);
        s.step((int64_t)chunk, dt, use_ba ? &bap : nullptr, use_th ? &thp : nullptr);
        done += chunk;
    }
    s.download(state);
    writer.push(frame, StateToSave::from(state));
    writer.finish();
    md_stats stt = s.stats();
    std::fprintf(stderr, "Calculated. steps=%lld rebuilds=%lld kernel_launches=%lld\n", (long long)stt.steps,
                 (long long)stt.rebuilds, (long long)stt.kernel_launches);
}

// cli/src/commands.rs:204-274 + core/src/save_data.rs:296-306
void cmd_solve_macro(const std::string &dir, Args &a)
{
    bool k = false, p = false, t = false, T = false, P = false, use_potentials = false;
    while (!a.done()) {
        std::string o = a.next("option");
        if (is(o, "-k", "--kinetic-energy")) k = true;
        else if (is(o, "-p", "--potential-energy")) p = true;
        else if (is(o, "-t", "--thermal-energy")) t = true;
        else if (is(o, "-T", "--temperature")) T = true;
        else if (is(o, "-P", "--pressure")) P = true;
        else if (is(o, "-c", "--custom")) throw Error(MD_ERR_UNSUPPORTED, "custom macro parameter is todo!() in the reference");
        else if (is(o, "-C", "--custom-name")) a.next(o);
        else if (is(o, "-A", "--all")) k = p = t = T = P = true;
        else if (is(o, nullptr, "--use-potentials")) use_potentials = true;
        else usage("unknown option for solve-macro-parameters: " + o);
    }
    PotentialsDatabase db;
    if (use_potentials) db.load_potentials_from_file(dir);
    ParticleDatabase::load_particles_data(dir);
    size_t end = last_frame(dir);
    std::ofstream f(dir + "/macro.csv", std::ios::trunc);
    f << "iteration,kinetic_energy,potential_energy,thermal_energy,unit_kinetic_energy,unit_potential_energy,"
         "unit_thermal_energy,temperature,pressure,custom\n";
    Session s(0, false);
    for (size_t i = 0; i <= end; ++i) {
        State state = StateToSave::load_from_file(dir, i).into_state();
        double n = (double)state.count();
        s.set_potentials(db, state);
        s.upload(state, false);
        s.update_force();  // forces are not stored in frames (commands.rs:234)
        md_macro_out m = s.macro(0);  // commands.rs:237-264: particle type 0
        double ke = k ? m.kinetic_energy : 0.0, pe = p ? m.potential_energy : 0.0;
        double te = (t || T) ? m.thermal_energy : 0.0;
        f << i << ',' << format_f64(ke) << ',' << format_f64(pe) << ',' << format_f64(te) << ',' << format_f64(ke / n)
          << ',' << format_f64(pe / n) << ',' << format_f64(te / n) << ',' << format_f64(T ? m.temperature : 0.0) << ','
          << format_f64(P ? m.pressure : 0.0) << ',' << format_f64(0.0) << '\n';
    }
}

void print_momentum(const std::string &dir, size_t frame, const char *title)  // commands.rs:276-305
{
    std::printf("%s\n", title);
    State state = StateToSave::load_from_file(dir, frame).into_state();
    for (size_t t = 0; t < state.particles.size(); ++t) {
        double p[3] = {0, 0, 0};
        for (auto &q : state.particles[t])
            for (int d = 0; d < 3; ++d) p[d] += q.velocity[d] * q.mass;
        std::printf("type = %zu;|p| = %.15f;p = [%.15f, %.15f, %.15f]\n", t,
                    std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]), p[0], p[1], p[2]);
    }
}

}  // namespace

int main(int argc, char **argv)
{
    std::string dir;
    bool time_it = false;
    size_t frames_per_save = 1;
    Args a;
    for (int i = 1; i < argc; ++i) a.v.push_back(argv[i]);
    std::string command;
    while (!a.done()) {
        std::string o = a.next("option");
        if (is(o, "-f", "--file")) dir = a.next(o);
        else if (o == "--time") time_it = true;
        else if (o == "--frames-per-save") frames_per_save = std::stoul(a.next(o));
        else { command = o; break; }
    }
    if (dir.empty()) usage("-f/--file is required");
    if (command.empty()) usage("missing command");
    auto start = std::chrono::steady_clock::now();
    try {
        if (command == "initialize") cmd_initialize(dir, a);
        else if (command == "solve") cmd_solve(dir, a, frames_per_save);
        else if (command == "solve-macro-parameters") cmd_solve_macro(dir, a);
        else if (command == "check-impulse") {
            ParticleDatabase::load_particles_data(dir);
            print_momentum(dir, 0, "First frame");
            print_momentum(dir, last_frame(dir), "Last frame");
        } else if (command == "copy-frames") {
            // extension (no GPU needed): frame -s is written again as the following -c frames through the asynchronous
            // frame writer `solve` uses — exercises the writer thread, its back-pressure and the cached bb.csv
            size_t from = 0, count = 0;
            while (!a.done()) {
                std::string o = a.next("option");
                if (is(o, "-s", "--state-number")) from = std::stoul(a.next(o));
                else if (is(o, "-c", "--iteration-count")) count = std::stoul(a.next(o));
                else usage("unknown option for copy-frames: " + o);
            }
            StateToSave data = StateToSave::load_from_file(dir, from);
            FrameWriter writer(dir, 2);
            for (size_t k = 1; k <= count; ++k) {
                StateToSave copy = data;
                copy.boundary_box[0] = data.boundary_box[0] + (double)k;  // every row of bb.csv distinguishable
                writer.push(from + k, std::move(copy));
            }
            writer.finish();
        } else if (command == "particle-count") {
            ParticleDatabase::load_particles_data(dir);
            std::printf("Particle count: %zu\n", StateToSave::load_from_file(dir, 0).into_state().count());
        } else if (command == "generate-default-potentials") {  // commands.rs:21-25
            PotentialsDatabase db;
            db.set_potential(0, 0, Potential::new_lennard_jones(0.3418, 1.712));
            db.save_potentials_to_file(dir);
        } else if (command == "set-potential") {  // commands.rs:27-41
            std::vector<std::string> types, params;
            std::string pot;
            while (!a.done()) {
                std::string o = a.next("option");
                if (is(o, "-i", "--particle-types")) types = a.values(o, 2, 2);
                else if (is(o, "-p", "--potential")) pot = a.next(o);
                else if (o == "--params") params = a.values(o, 1, 16);
                else usage("unknown option for set-potential: " + o);
            }
            if (pot != "lennard-jones") throw Error(MD_ERR_UNSUPPORTED, "PotentialChoose::Custom is todo!()");
            if (types.size() != 2 || params.size() < 2) usage("set-potential needs -i a b -p lennard-jones --params s e");
            PotentialsDatabase db;
            db.load_potentials_from_file(dir);
            db.set_potential((uint16_t)std::stoul(types[0]), (uint16_t)std::stoul(types[1]),
                             Potential::new_lennard_jones(std::stod(params[0]), std::stod(params[1])));
            db.save_potentials_to_file(dir);
        } else if (command == "generate-velocities-histogram") {  // commands.rs:318-347
            size_t frame = 0;
            std::vector<std::string> types;
            while (!a.done()) {
                std::string o = a.next("option");
                if (is(o, "-s", "--state-number")) frame = std::stoul(a.next(o));
                else if (o == "--particle-types") types = a.values(o, 1, 65535);
                else usage("unknown option for generate-velocities-histogram: " + o);
            }
            ParticleDatabase::load_particles_data(dir);
            State state = StateToSave::load_from_file(dir, frame).into_state();
            std::ofstream f(dir + "/hist.csv", std::ios::trunc);
            f << "x,y,z\n";
            for (auto &ts : types)
                for (auto &q : state.particles.at(std::stoul(ts)))
                    f << format_f64(q.velocity[0]) << ',' << format_f64(q.velocity[1]) << ','
                      << format_f64(q.velocity[2]) << '\n';
        } else usage("unknown command " + command);
    } catch (const Error &e) {
        std::cerr << "moldyn_cli: " << e.what() << " (code " << e.code << ")\n";
        return 1;
    } catch (const std::exception &e) {
        std::cerr << "moldyn_cli: " << e.what() << "\n";
        return 1;
    }
    if (time_it) {
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
        std::printf("Time elapsed: %.9g\n", secs);  // cli/src/main.rs:99-102
    }
    return 0;
}
