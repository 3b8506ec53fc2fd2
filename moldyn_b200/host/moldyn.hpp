// moldyn.hpp — C++ host-side mirror of the reference's solver/core interface for the `solve` path.
//
// The reference is compiled (Rust) code and this image has no Rust toolchain, so the host layer above the C ABI
// (include/moldyn_b200.h) is C++.  Names, argument meaning and error behaviour follow the reference:
//   Particle, State, ParticleDatabase            core/src/particle.rs, core/src/particles_database.rs
//   ParticleToSave, StateToSave (CSV frames)      core/src/save_data.rs
//   Potential, PotentialsDatabase, update_force   solver/src/solver/potential.rs
//   Integrator, Thermostat, Barostat              solver/src/solver/integrator.rs, solver/src/initializer/*.rs
//   macro_parameters::get_*                       solver/src/macro_parameters/*.rs
//   initialize_particles(_position), velocities   solver/src/initializer/{position,velocity}.rs  (input generator)
// All physics of the step runs in the CUDA library; the only host arithmetic here is I/O, marshalling and the
// one-shot O(N) input generator.  Where the reference panics (expect/unwrap/todo!), this layer throws moldyn::Error.
#pragma once

#include <array>
#include <cstdint>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/moldyn_b200.h"

namespace moldyn {

constexpr double K_B = 1.380648528;  // core/src/lib.rs:15

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

using Vector3 = std::array<double, 3>;

// core/src/particle.rs:6-23
struct Particle {
    Vector3 position{0, 0, 0}, velocity{0, 0, 0}, force{0, 0, 0};
    double potential = 0.0;
    double temp = 0.0;  // Σ F(i,j)·r(i,j)
    double mass = 1.0;
    double radius = 0.1;
    uint16_t id = 0;
};

// core/src/particles_database.rs:52-237 — process-global id → (name, mass, radius); db.csv
struct ParticleData {
    std::string name;
    double mass, radius;
};
class ParticleDatabase {
public:
    static void add(uint16_t id, const std::string &name, double mass, double radius);
    static std::optional<double> get_particle_mass(uint16_t id);
    static std::optional<double> get_particle_radius(uint16_t id);
    static std::optional<std::string> get_particle_name(uint16_t id);
    static void clear_particles();
    static void save_particles_data(const std::string &dir);  // <dir>/db.csv  (id,name,mass,radius)
    static void load_particles_data(const std::string &dir);

private:
    static std::map<uint16_t, ParticleData> &data();
};

// core/src/particle.rs:25-32, 118-142
struct State {
    std::vector<std::vector<Particle>> particles;  // indexed by type id
    Vector3 boundary_box{0, 0, 0};
    void apply_boundary_conditions();  // runs on the device through a zero-length step? No: plain wrap, see .cpp
    size_t count() const;
};

// core/src/save_data.rs:11-27, 79-230
struct ParticleToSave {
    uint16_t id;
    double position_x, position_y, position_z, velocity_x, velocity_y, velocity_z;
};
struct StateToSave {
    std::vector<ParticleToSave> particles;
    Vector3 boundary_box{0, 0, 0};
    static StateToSave from(const State &state);                                // save_data.rs:113-127
    State into_state() const;                                                   // save_data.rs:129-151
    void save_to_file(const std::string &dir, size_t state_number) const;       // save_data.rs:191-213
    static StateToSave load_from_file(const std::string &dir, size_t state_number);  // save_data.rs:215-230
};

// Frame output off the critical path: `solve` hands every downloaded frame to a writer thread and goes on
// stepping on the GPU while the CSV text is produced; at most `depth` frames wait in memory.  The frames on disk are exactly
// what StateToSave::save_to_file writes (save_data.rs:191-213); bb.csv is read once and then kept in memory instead of being
// re-read for every frame (the reference's Θ(frames²) rewrite, save_data.rs:165-184).
class FrameWriter {
public:
    FrameWriter(std::string dir, size_t depth = 2);
    ~FrameWriter();                                   // joins; errors of the writer thread are dropped here
    void push(size_t state_number, StateToSave frame);  // blocks while `depth` frames are pending; rethrows writer errors
    void finish();                                    // waits for everything to be on disk; rethrows writer errors
private:
    struct Impl;
    Impl *impl_;
};

// Shortest round-trip decimal in the layout of the `ryu` crate the reference's csv writer uses ("1.0", "1e-7").
std::string format_f64(double v);

// solver/src/solver/potential.rs:12-87
struct Potential {
    double sigma, eps, r_cut, u_cut;
    static Potential new_lennard_jones(double sigma, double eps);
    std::pair<double, double> get_potential_and_force(double r) const;
    double get_radius_cut() const { return r_cut; }
};

// solver/src/solver/potential.rs:89-155 — potentials.json: {"0,0": {"LennardJones": {sigma, eps, r_cut, u_cut}}}
class PotentialsDatabase {
public:
    PotentialsDatabase();
    void set_potential(uint16_t id0, uint16_t id1, const Potential &p);
    const Potential &get_potential(uint16_t id0, uint16_t id1) const;
    void save_potentials_to_file(const std::string &dir) const;
    void load_potentials_from_file(const std::string &dir);

private:
    std::map<std::pair<uint16_t, uint16_t>, Potential> potentials_;
    Potential default_potential_;
};

// solver/src/initializer/thermostat.rs:4-22 / barostat.rs:4-19 (Custom variants are todo!() in the reference)
struct Thermostat {
    enum Kind { Berendsen = MD_THERMOSTAT_BERENDSEN, NoseHoover = MD_THERMOSTAT_NOSE_HOOVER, Custom = 99 } kind;
    double tau = 1.0, lambda = 0.0, psi = 0.0;
};
struct Barostat {
    enum Kind { Berendsen = MD_BAROSTAT_BERENDSEN, Custom = 99 } kind;
    double beta = 1.0, tau = 1.0, myu = 0.0;
};

// Device-resident session over one md_ctx (the fast path: upload once, step many, download at frame boundaries).
class Session {
public:
    explicit Session(int device = 0, bool exact = false, double skin = 0.0);
    ~Session();
    Session(const Session &) = delete;
    Session &operator=(const Session &) = delete;

    void set_potential(const Potential &p);
    // every PotentialsDatabase entry the State's particle types can meet (one type: its own pair)
    void set_potentials(const PotentialsDatabase &db, const State &state);
    // false (default): the reference's one-sided cross-type accumulation (potential.rs:168-176); true: symmetric table
    void set_symmetric_cross_type_forces(bool symmetric);
    // State.particles type by type (one type: md_upload_state; several: md_upload_state_typed)
    void upload(const State &state, bool with_forces);
    void download(State &state);
    void update_force();
    void step(int64_t n_steps, double dt, std::pair<Barostat *, double> *barostat,
              std::pair<Thermostat *, double> *thermostat);
    md_macro_out macro();
    md_macro_out macro(uint16_t particle_type_id);
    md_stats stats();
    md_ctx *raw() { return ctx_; }

private:
    void check(int rc);
    md_ctx *ctx_ = nullptr;
    std::vector<double> pos_, vel_, force_, pot_, vir_;
    std::vector<uint16_t> types_;  // type ids of the uploaded State, in upload order
};

// solver/src/solver/potential.rs:158 — per-call semantics (upload → forces → download) on a shared session
void update_force(const PotentialsDatabase &db, State &state);

// solver/src/solver/integrator.rs:5-15
struct Integrator {
    enum Kind { VerletMethod, Custom } kind = VerletMethod;
    void calculate(const PotentialsDatabase &db, State &state, double delta_time,
                   std::optional<std::pair<Barostat *, double>> &barostat,
                   std::optional<std::pair<Thermostat *, double>> &thermostat) const;
};

// solver/src/macro_parameters/*.rs — evaluated on the device for particle type `particle_type_id`
namespace macro_parameters {
Vector3 get_center_of_mass_velocity(const State &state, uint16_t particle_type_id);
Vector3 get_momentum_of_system(const State &state, uint16_t particle_type_id);
double get_kinetic_energy(const State &state, uint16_t particle_type_id);
double get_thermal_energy(const State &state, uint16_t particle_type_id, const Vector3 &vcom);
double get_potential_energy(const State &state, uint16_t particle_type_id);
double get_temperature(double thermal_energy, size_t number_particles);
double get_pressure(const State &state, uint16_t particle_type_id, const Vector3 &vcom);
}  // namespace macro_parameters

// solver/src/initializer/position.rs, velocity.rs — one-shot host-side input generator
namespace initializer {
enum class InitError { ParticleIdDidNotFound, TooBig, OutOfBoundary };
enum class UnitCell { U, FCC };
State initialize_particles(const std::vector<size_t> &number_particles, const Vector3 &boundary);
void initialize_particles_position(UnitCell cell, State &state, uint16_t particle_id, const Vector3 &start,
                                   const std::array<size_t, 3> &grid_size, double unit_cell_size);
void initialize_velocities_maxwell_boltzmann(State &state, double temperature, uint16_t particle_id,
                                             uint64_t seed = 0 /* 0 = nondeterministic like thread_rng */);
struct InitException : std::runtime_error {
    InitError error;
    InitException(InitError e, const char *m) : std::runtime_error(m), error(e) {}
};
}  // namespace initializer

}  // namespace moldyn
