// moldyn.cpp — implementation of the host-side mirror (see moldyn.hpp).  Frame / database / potentials file
// formats follow the reference byte layout: core/src/save_data.rs:153-230, core/src/particles_database.rs:157-228,
// solver/src/solver/potential.rs:104-139.
#include "moldyn.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <charconv>
#include <condition_variable>
#include <deque>
#include <exception>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <random>
#include <sstream>

namespace moldyn {

namespace {

void make_dirs(const std::string &path)
{
    std::string cur;
    for (size_t i = 0; i <= path.size(); ++i) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty()) mkdir(cur.c_str(), 0777);
        }
        if (i < path.size()) cur.push_back(path[i]);
    }
}

bool file_exists(const std::string &p)
{
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

std::vector<std::string> split_csv_line(const std::string &line)
{
    std::vector<std::string> out;
    std::string cur;
    bool quoted = false;
    for (size_t i = 0; i < line.size(); ++i) {
        char c = line[i];
        if (quoted) {
            if (c == '"' && i + 1 < line.size() && line[i + 1] == '"') { cur.push_back('"'); ++i; }
            else if (c == '"') quoted = false;
            else cur.push_back(c);
        } else if (c == '"') quoted = true;
        else if (c == ',') { out.push_back(cur); cur.clear(); }
        else if (c != '\r') cur.push_back(c);
    }
    out.push_back(cur);
    return out;
}

double parse_f64(const std::string &s)
{
    char *end = nullptr;
    double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str()) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't parse row: bad float '" + s + "'");
    return v;
}

std::string csv_field(const std::string &s)
{
    if (s.find_first_of(",\"\n") == std::string::npos) return s;
    std::string o = "\"";
    for (char c : s) { if (c == '"') o.push_back('"'); o.push_back(c); }
    return o + "\"";
}

}  // namespace

// ---- shortest round-trip float formatting in ryu's layout ------------------------------------------------
std::string format_f64(double v)
{
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);  // shortest digits
    std::string s(buf, res.ptr);
    bool neg = s[0] == '-';
    if (neg) s.erase(0, 1);
    size_t epos = s.find('e');
    std::string mant = s.substr(0, epos);
    int exp10 = std::atoi(s.c_str() + epos + 1);
    std::string digits;
    for (char c : mant) if (c != '.') digits.push_back(c);
    int len = (int)digits.size();
    int k = exp10 - (len - 1);  // value = digits * 10^k
    int kk = len + k;           // position of the decimal point
    std::string out;
    if (0 <= k && kk <= 16) {
        out = digits + std::string((size_t)k, '0') + ".0";
    } else if (0 < kk && kk <= 16) {
        out = digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
    } else if (-5 < kk && kk <= 0) {
        out = "0." + std::string((size_t)(-kk), '0') + digits;
    } else if (len == 1) {
        out = digits + "e" + std::to_string(kk - 1);
    } else {
        out = digits.substr(0, 1) + "." + digits.substr(1) + "e" + std::to_string(kk - 1);
    }
    return neg ? "-" + out : out;
}

// ---- ParticleDatabase -------------------------------------------------------------------------------------
std::map<uint16_t, ParticleData> &ParticleDatabase::data()
{
    static std::map<uint16_t, ParticleData> d;
    return d;
}
void ParticleDatabase::add(uint16_t id, const std::string &name, double mass, double radius)
{
    data()[id] = ParticleData{name, mass, radius};
}
std::optional<double> ParticleDatabase::get_particle_mass(uint16_t id)
{
    auto it = data().find(id);
    return it == data().end() ? std::nullopt : std::optional<double>(it->second.mass);
}
std::optional<double> ParticleDatabase::get_particle_radius(uint16_t id)
{
    auto it = data().find(id);
    return it == data().end() ? std::nullopt : std::optional<double>(it->second.radius);
}
std::optional<std::string> ParticleDatabase::get_particle_name(uint16_t id)
{
    auto it = data().find(id);
    return it == data().end() ? std::nullopt : std::optional<std::string>(it->second.name);
}
void ParticleDatabase::clear_particles() { data().clear(); }

void ParticleDatabase::save_particles_data(const std::string &dir)
{
    make_dirs(dir);
    std::ofstream f(dir + "/db.csv", std::ios::trunc);
    if (!f) throw Error(MD_ERR_INVALID_ARGUMENT, "CantOpen: " + dir + "/db.csv");
    f << "id,name,mass,radius\n";
    for (auto &kv : data())
        f << kv.first << ',' << csv_field(kv.second.name) << ',' << format_f64(kv.second.mass) << ','
          << format_f64(kv.second.radius) << '\n';
}

void ParticleDatabase::load_particles_data(const std::string &dir)
{
    std::ifstream f(dir + "/db.csv");
    if (!f) throw Error(MD_ERR_INVALID_ARGUMENT, "CantOpen: " + dir + "/db.csv");
    std::string line;
    std::getline(f, line);  // header
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        auto c = split_csv_line(line);
        if (c.size() < 4) throw Error(MD_ERR_INVALID_ARGUMENT, "CantRead: db.csv row");
        uint16_t id = (uint16_t)std::stoul(c[0]);
        data().emplace(id, ParticleData{c[1], parse_f64(c[2]), parse_f64(c[3])});  // entry().or_insert()
    }
}

// ---- State ---------------------------------------------------------------------------------------------------
size_t State::count() const
{
    size_t n = 0;
    for (auto &t : particles) n += t.size();
    return n;
}

// core/src/particle.rs:120-142 as a data-model helper; inside a step the wrap runs in k_kick_drift on the device.
void State::apply_boundary_conditions()
{
    for (auto &t : particles)
        for (auto &p : t)
            for (int d = 0; d < 3; ++d) {
                if (p.position[d] < 0.0) p.position[d] += boundary_box[d];
                else if (p.position[d] >= boundary_box[d]) p.position[d] -= boundary_box[d];
            }
}

// ---- StateToSave ---------------------------------------------------------------------------------------------
StateToSave StateToSave::from(const State &state)
{
    StateToSave s;
    s.boundary_box = state.boundary_box;
    s.particles.reserve(state.count());
    for (auto &t : state.particles)
        for (auto &p : t)
            s.particles.push_back({p.id, p.position[0], p.position[1], p.position[2], p.velocity[0], p.velocity[1],
                                   p.velocity[2]});
    return s;
}

State StateToSave::into_state() const
{
    State st;
    uint16_t max_id = 0;
    for (auto &p : particles) max_id = std::max(max_id, p.id);
    st.particles.resize((size_t)max_id + 1);
    for (auto &p : particles) {
        auto mass = ParticleDatabase::get_particle_mass(p.id);
        if (!mass) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't convert particle");  // save_data.rs:145
        Particle q;
        q.position = {p.position_x, p.position_y, p.position_z};
        q.velocity = {p.velocity_x, p.velocity_y, p.velocity_z};
        q.mass = *mass;
        q.radius = *ParticleDatabase::get_particle_radius(p.id);
        q.id = p.id;
        st.particles[p.id].push_back(q);
    }
    st.boundary_box = boundary_box;
    return st;
}

static std::vector<Vector3> get_bbs(const std::string &path)
{
    std::vector<Vector3> bbs;
    std::ifstream f(path);
    if (!f) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't open file " + path);
    std::string line;
    std::getline(f, line);
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        auto c = split_csv_line(line);
        if (c.size() < 3) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't deserialize bb.csv");
        bbs.push_back({parse_f64(c[0]), parse_f64(c[1]), parse_f64(c[2])});
    }
    return bbs;
}

// data/<n>.csv of one frame (save_data.rs:186-213)
static void write_frame_csv(const StateToSave &st, const std::string &dir, size_t state_number)
{
    make_dirs(dir + "/data");
    const std::string path = dir + "/data/" + std::to_string(state_number) + ".csv";
    std::unique_ptr<FILE, int (*)(FILE *)> f(std::fopen(path.c_str(), "w"), std::fclose);
    if (!f) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't write to file " + path);
    std::string buf;
    buf.reserve(st.particles.size() * 140 + 128);
    buf += "id,position_x,position_y,position_z,velocity_x,velocity_y,velocity_z\n";
    for (auto &p : st.particles) {
        buf += std::to_string(p.id);
        for (double v : {p.position_x, p.position_y, p.position_z, p.velocity_x, p.velocity_y, p.velocity_z}) {
            buf.push_back(',');
            buf += format_f64(v);
        }
        buf.push_back('\n');
    }
    if (std::fwrite(buf.data(), 1, buf.size(), f.get()) != buf.size())
        throw Error(MD_ERR_INVALID_ARGUMENT, "Can't write");
}

// bb.csv: row n = boundary box of frame n (save_data.rs:165-184).  `bbs` is the file's content as known to the caller; the
// reference re-reads and re-writes the whole file per frame, appending is equivalent when the frame is the next row.
static void write_bb_row(std::vector<Vector3> &bbs, const std::string &dir, size_t state_number, const Vector3 &box)
{
    const std::string bb_path = dir + "/bb.csv";
    if (bbs.size() > state_number) {
        bbs[state_number] = box;
        std::ofstream f(bb_path, std::ios::trunc);
        f << "x,y,z\n";
        for (auto &b : bbs) f << format_f64(b[0]) << ',' << format_f64(b[1]) << ',' << format_f64(b[2]) << '\n';
    } else {
        const bool fresh = !file_exists(bb_path) || bbs.empty();
        std::ofstream f(bb_path, fresh ? std::ios::trunc : std::ios::app);
        if (fresh) f << "x,y,z\n";
        f << format_f64(box[0]) << ',' << format_f64(box[1]) << ',' << format_f64(box[2]) << '\n';
        bbs.push_back(box);
    }
}

void StateToSave::save_to_file(const std::string &dir, size_t state_number) const
{
    make_dirs(dir);
    const std::string bb_path = dir + "/bb.csv";
    std::vector<Vector3> bbs;
    if (file_exists(bb_path)) bbs = get_bbs(bb_path);
    write_bb_row(bbs, dir, state_number, boundary_box);
    write_frame_csv(*this, dir, state_number);
}

struct FrameWriter::Impl {
    std::string dir;
    size_t depth;
    std::vector<Vector3> bbs;
    std::deque<std::pair<size_t, StateToSave>> queue;
    std::mutex m;
    std::condition_variable cv_push, cv_pop;
    bool closing = false, busy = false;
    std::exception_ptr error;
    std::thread worker;

    void run()
    {
        for (;;) {
            std::pair<size_t, StateToSave> item;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_pop.wait(lk, [&] { return closing || !queue.empty(); });
                if (queue.empty()) return;
                item = std::move(queue.front());
                queue.pop_front();
                busy = true;
            }
            try {
                write_bb_row(bbs, dir, item.first, item.second.boundary_box);
                write_frame_csv(item.second, dir, item.first);
            } catch (...) {
                std::lock_guard<std::mutex> lk(m);
                if (!error) error = std::current_exception();
            }
            {
                std::lock_guard<std::mutex> lk(m);
                busy = false;
            }
            cv_push.notify_all();
        }
    }
};

FrameWriter::FrameWriter(std::string dir, size_t depth) : impl_(new Impl)
{
    impl_->dir = std::move(dir);
    impl_->depth = depth ? depth : 1;
    make_dirs(impl_->dir);
    const std::string bb_path = impl_->dir + "/bb.csv";
    if (file_exists(bb_path)) impl_->bbs = get_bbs(bb_path);
    impl_->worker = std::thread([this] { impl_->run(); });
}

FrameWriter::~FrameWriter()
{
    {
        std::lock_guard<std::mutex> lk(impl_->m);
        impl_->closing = true;
    }
    impl_->cv_pop.notify_all();
    if (impl_->worker.joinable()) impl_->worker.join();
    delete impl_;
}

void FrameWriter::push(size_t state_number, StateToSave frame)
{
    std::unique_lock<std::mutex> lk(impl_->m);
    impl_->cv_push.wait(lk, [&] { return impl_->error || impl_->queue.size() < impl_->depth; });
    if (impl_->error) std::rethrow_exception(impl_->error);
    impl_->queue.emplace_back(state_number, std::move(frame));
    lk.unlock();
    impl_->cv_pop.notify_all();
}

void FrameWriter::finish()
{
    std::unique_lock<std::mutex> lk(impl_->m);
    impl_->cv_push.wait(lk, [&] { return impl_->queue.empty() && !impl_->busy; });
    if (impl_->error) std::rethrow_exception(impl_->error);
}

StateToSave StateToSave::load_from_file(const std::string &dir, size_t state_number)
{
    StateToSave s;
    auto bbs = get_bbs(dir + "/bb.csv");
    if (state_number >= bbs.size()) throw Error(MD_ERR_INVALID_ARGUMENT, "bb.csv has no row for this frame");
    s.boundary_box = bbs[state_number];
    const std::string path = dir + "/data/" + std::to_string(state_number) + ".csv";
    std::ifstream f(path);
    if (!f) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't open file " + path);
    std::string line;
    std::getline(f, line);
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        auto c = split_csv_line(line);
        if (c.size() < 7) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't parse row");
        s.particles.push_back({(uint16_t)std::stoul(c[0]), parse_f64(c[1]), parse_f64(c[2]), parse_f64(c[3]),
                               parse_f64(c[4]), parse_f64(c[5]), parse_f64(c[6])});
    }
    return s;
}

// ---- Potential / PotentialsDatabase -----------------------------------------------------------------------------
Potential Potential::new_lennard_jones(double sigma, double eps)
{
    Potential p{sigma, eps, 0.0, 0.0};
    md_lj_new(sigma, eps, &p.r_cut, &p.u_cut);
    return p;
}

std::pair<double, double> Potential::get_potential_and_force(double r) const
{
    double u, f;
    md_lj_potential_and_force(sigma, eps, r_cut, u_cut, r, &u, &f);
    return {u, f};
}

PotentialsDatabase::PotentialsDatabase() : default_potential_(Potential::new_lennard_jones(0.3418, 1.712)) {}

void PotentialsDatabase::set_potential(uint16_t id0, uint16_t id1, const Potential &p)
{
    potentials_[{std::min(id0, id1), std::max(id0, id1)}] = p;
}

const Potential &PotentialsDatabase::get_potential(uint16_t id0, uint16_t id1) const
{
    auto it = potentials_.find({std::min(id0, id1), std::max(id0, id1)});
    return it == potentials_.end() ? default_potential_ : it->second;
}

void PotentialsDatabase::save_potentials_to_file(const std::string &dir) const
{
    make_dirs(dir);
    std::ofstream f(dir + "/potentials.json", std::ios::trunc);
    if (!f) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't create file");
    f << "{";
    bool first = true;
    for (auto &kv : potentials_) {
        f << (first ? "\n" : ",\n") << "  \"" << kv.first.first << "," << kv.first.second << "\": {\n"
          << "    \"LennardJones\": {\n"
          << "      \"sigma\": " << format_f64(kv.second.sigma) << ",\n"
          << "      \"eps\": " << format_f64(kv.second.eps) << ",\n"
          << "      \"r_cut\": " << format_f64(kv.second.r_cut) << ",\n"
          << "      \"u_cut\": " << format_f64(kv.second.u_cut) << "\n    }\n  }";
        first = false;
    }
    f << "\n}";
}

namespace {
// Minimal reader for the potentials.json shape the reference writes (serde_json pretty, externally tagged enum).
struct JsonCursor {
    const std::string &s;
    size_t i = 0;
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
    void expect(char c) { if (!eat(c)) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't load data from file: bad potentials.json"); }
    std::string str()
    {
        expect('"');
        std::string o;
        while (i < s.size() && s[i] != '"') o.push_back(s[i++]);
        expect('"');
        return o;
    }
    double num()
    {
        ws();
        char *end = nullptr;
        double v = std::strtod(s.c_str() + i, &end);
        if (end == s.c_str() + i) throw Error(MD_ERR_INVALID_ARGUMENT, "bad number in potentials.json");
        i = (size_t)(end - s.c_str());
        return v;
    }
};
}  // namespace

void PotentialsDatabase::load_potentials_from_file(const std::string &dir)
{
    std::ifstream f(dir + "/potentials.json");
    if (!f) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't open file");
    std::stringstream ss;
    ss << f.rdbuf();
    std::string text = ss.str();
    JsonCursor c{text};
    c.expect('{');
    if (c.eat('}')) return;
    do {
        std::string key = c.str();
        size_t comma = key.find(',');
        if (comma == std::string::npos) throw Error(MD_ERR_INVALID_ARGUMENT, "Can't convert " + key + " to i16");
        uint16_t a = (uint16_t)std::stoul(key.substr(0, comma)), b = (uint16_t)std::stoul(key.substr(comma + 1));
        c.expect(':');
        c.expect('{');
        std::string tag = c.str();
        if (tag != "LennardJones") throw Error(MD_ERR_UNSUPPORTED, "Potential::Custom is todo!() in the reference");
        c.expect(':');
        c.expect('{');
        Potential p{0, 0, 0, 0};
        do {
            std::string field = c.str();
            c.expect(':');
            double v = c.num();
            if (field == "sigma") p.sigma = v;
            else if (field == "eps") p.eps = v;
            else if (field == "r_cut") p.r_cut = v;
            else if (field == "u_cut") p.u_cut = v;
        } while (c.eat(','));
        c.expect('}');
        c.expect('}');
        potentials_[{a, b}] = p;  // inserted with the key as written (potential.rs:137)
    } while (c.eat(','));
    c.expect('}');
}

// ---- Session -----------------------------------------------------------------------------------------------------
namespace {
// The particle types of a State that have atoms.  One such type (any id) is the single-type device path; several must be the
// ids 0..T-1 without a gap: the reference indexes particle_type[0] of every entry of State.particles and panics on an empty one
// (integrator.rs:29).
std::vector<uint16_t> types_of(const State &state)
{
    std::vector<uint16_t> t;
    for (size_t k = 0; k < state.particles.size(); ++k)
        if (!state.particles[k].empty()) t.push_back((uint16_t)k);
    if (t.empty()) throw Error(MD_ERR_INVALID_ARGUMENT, "empty State (the reference panics on particle_type[0])");
    if (t.size() > 1 && t.size() != state.particles.size())
        throw Error(MD_ERR_INVALID_ARGUMENT, "a particle type without atoms between the others (the reference panics on "
                                             "particle_type[0], integrator.rs:29)");
    return t;
}
}  // namespace

Session::Session(int device, bool exact, double skin)
{
    md_config cfg{};
    cfg.device = device;
    cfg.force_mode = exact ? MD_FORCE_EXACT : MD_FORCE_FAST;
    cfg.skin = skin;
    int rc = md_create(&cfg, &ctx_);
    if (rc != MD_OK) throw Error(rc, md_last_error(nullptr));
}

Session::~Session() { md_destroy(ctx_); }

void Session::check(int rc)
{
    if (rc != MD_OK) throw Error(rc, md_last_error(ctx_));
}

void Session::set_potential(const Potential &p) { check(md_set_potential_lj(ctx_, p.sigma, p.eps, p.r_cut, p.u_cut)); }

void Session::set_potentials(const PotentialsDatabase &db, const State &state)
{
    const auto types = types_of(state);
    if (types.size() == 1) {
        set_potential(db.get_potential(types[0], types[0]));
        return;
    }
    for (uint16_t a : types)
        for (uint16_t b : types)
            if (a <= b) {
                const Potential &p = db.get_potential(a, b);  // (min, max) key or the default (potential.rs:147-155)
                check(md_set_potential_pair(ctx_, a, b, p.sigma, p.eps, p.r_cut, p.u_cut));
            }
}

void Session::set_symmetric_cross_type_forces(bool symmetric)
{
    check(md_set_cross_type_mode(ctx_, symmetric ? MD_CROSS_SYMMETRIC : MD_CROSS_REFERENCE));
}

void Session::upload(const State &state, bool with_forces)
{
    types_ = types_of(state);
    size_t n = 0;
    for (uint16_t t : types_) n += state.particles[t].size();
    pos_.resize(3 * n); vel_.resize(3 * n); force_.resize(3 * n); pot_.resize(n); vir_.resize(n);
    std::vector<int64_t> counts;
    std::vector<double> masses;
    size_t i = 0;
    for (uint16_t t : types_) {  // State.particles flattened type by type
        const auto &ps = state.particles[t];
        counts.push_back((int64_t)ps.size());
        masses.push_back(ps[0].mass);  // integrator.rs:30: the mass of the type's first particle
        for (const Particle &q : ps) {
            for (int d = 0; d < 3; ++d) {
                pos_[3 * i + d] = q.position[d];
                vel_[3 * i + d] = q.velocity[d];
                force_[3 * i + d] = q.force[d];
            }
            pot_[i] = q.potential;
            vir_[i] = q.temp;
            ++i;
        }
    }
    check(md_upload_state_typed(ctx_, (int64_t)n, pos_.data(), vel_.data(), with_forces ? force_.data() : nullptr,
                                with_forces ? pot_.data() : nullptr, with_forces ? vir_.data() : nullptr,
                                (int32_t)counts.size(), counts.data(), masses.data(), state.boundary_box.data()));
}

void Session::download(State &state)
{
    if (types_ != types_of(state)) throw Error(MD_ERR_INVALID_ARGUMENT, "download into a State of another shape than the uploaded one");
    size_t n = 0;
    for (uint16_t t : types_) n += state.particles[t].size();
    pos_.resize(3 * n); vel_.resize(3 * n); force_.resize(3 * n); pot_.resize(n); vir_.resize(n);
    check(md_download_state(ctx_, pos_.data(), vel_.data(), force_.data(), pot_.data(), vir_.data(),
                            state.boundary_box.data()));
    size_t i = 0;
    for (uint16_t t : types_)
        for (Particle &q : state.particles[t]) {
            for (int d = 0; d < 3; ++d) {
                q.position[d] = pos_[3 * i + d];
                q.velocity[d] = vel_[3 * i + d];
                q.force[d] = force_[3 * i + d];
            }
            q.potential = pot_[i];
            q.temp = vir_[i];
            ++i;
        }
}

void Session::update_force() { check(md_update_force(ctx_)); }

void Session::step(int64_t n_steps, double dt, std::pair<Barostat *, double> *barostat,
                   std::pair<Thermostat *, double> *thermostat)
{
    md_thermostat th{};
    md_barostat ba{};
    if (thermostat) {
        if (thermostat->first->kind == Thermostat::Custom)
            throw Error(MD_ERR_UNSUPPORTED, "Thermostat::Custom is todo!() in the reference");
        th.kind = (int)thermostat->first->kind;
        th.tau = thermostat->first->tau;
        th.target = thermostat->second;
        th.psi = thermostat->first->psi;
    }
    if (barostat) {
        if (barostat->first->kind == Barostat::Custom)
            throw Error(MD_ERR_UNSUPPORTED, "Barostat::Custom is todo!() in the reference");
        ba.kind = (int)barostat->first->kind;
        ba.beta = barostat->first->beta;
        ba.tau = barostat->first->tau;
        ba.target = barostat->second;
    }
    check(md_step(ctx_, n_steps, dt, thermostat ? &th : nullptr, barostat ? &ba : nullptr));
    if (thermostat) { thermostat->first->lambda = th.lambda; thermostat->first->psi = th.psi; }
    if (barostat) barostat->first->myu = ba.myu;
}

md_macro_out Session::macro()
{
    md_macro_out m{};
    check(md_macro(ctx_, &m));
    return m;
}

md_macro_out Session::macro(uint16_t particle_type_id)
{
    md_macro_out m{};
    if (types_.size() <= 1) check(md_macro(ctx_, &m));  // (the one type that has atoms, whatever its id)
    else check(md_macro_type(ctx_, particle_type_id, &m));
    return m;
}

md_stats Session::stats()
{
    md_stats s{};
    check(md_get_stats(ctx_, &s));
    return s;
}

// ---- per-call forms with the reference's signatures -------------------------------------------------------------------
namespace {
Session &shared_session()
{
    static Session s(0, std::getenv("MOLDYN_B200_EXACT") != nullptr);
    return s;
}
}  // namespace

void update_force(const PotentialsDatabase &db, State &state)
{
    Session &s = shared_session();
    s.set_potentials(db, state);
    s.upload(state, false);
    s.update_force();
    s.download(state);
}

void Integrator::calculate(const PotentialsDatabase &db, State &state, double delta_time,
                           std::optional<std::pair<Barostat *, double>> &barostat,
                           std::optional<std::pair<Thermostat *, double>> &thermostat) const
{
    if (kind != VerletMethod) throw Error(MD_ERR_UNSUPPORTED, "Integrator::Custom is todo!() in the reference");
    Session &s = shared_session();
    s.set_potentials(db, state);
    s.upload(state, true);
    s.step(1, delta_time, barostat ? &*barostat : nullptr, thermostat ? &*thermostat : nullptr);
    s.download(state);
}

namespace macro_parameters {
namespace {
md_macro_out macro_of(const State &state, uint16_t particle_type_id)
{
    Session &s = shared_session();
    s.upload(state, true);
    return s.macro(particle_type_id);
}
}  // namespace
Vector3 get_center_of_mass_velocity(const State &state, uint16_t t)
{
    auto m = macro_of(state, t);
    return {m.vcom[0], m.vcom[1], m.vcom[2]};
}
Vector3 get_momentum_of_system(const State &state, uint16_t t)
{
    auto m = macro_of(state, t);
    return {m.momentum[0], m.momentum[1], m.momentum[2]};
}
double get_kinetic_energy(const State &state, uint16_t t) { return macro_of(state, t).kinetic_energy; }
double get_thermal_energy(const State &state, uint16_t t, const Vector3 &) { return macro_of(state, t).thermal_energy; }
double get_potential_energy(const State &state, uint16_t t) { return macro_of(state, t).potential_energy; }
double get_temperature(double thermal_energy, size_t number_particles)
{
    double t = (2.0 * thermal_energy) / (3.0 * (double)number_particles * K_B);
    return t * 100.0;
}
double get_pressure(const State &state, uint16_t t, const Vector3 &) { return macro_of(state, t).pressure; }
}  // namespace macro_parameters

// ---- initializer (input generator) ----------------------------------------------------------------------------------------
namespace initializer {

State initialize_particles(const std::vector<size_t> &number_particles, const Vector3 &boundary)
{
    State st;
    for (size_t i = 0; i < number_particles.size(); ++i) {
        auto mass = ParticleDatabase::get_particle_mass((uint16_t)i);
        if (!mass) throw InitException(InitError::ParticleIdDidNotFound, "ParticleIdDidNotFound");
        Particle p;
        p.id = (uint16_t)i;
        p.mass = *mass;
        p.radius = *ParticleDatabase::get_particle_radius((uint16_t)i);
        st.particles.emplace_back(number_particles[i], p);
    }
    st.boundary_box = boundary;
    return st;
}

void initialize_particles_position(UnitCell cell, State &state, uint16_t particle_id, const Vector3 &start,
                                   const std::array<size_t, 3> &g, double l)
{
    if (!ParticleDatabase::get_particle_mass(particle_id))
        throw InitException(InitError::ParticleIdDidNotFound, "ParticleIdDidNotFound");
    if ((double)g[0] * l > state.boundary_box[0] || (double)g[1] * l > state.boundary_box[1] ||
        (double)g[2] * l > state.boundary_box[2])
        throw InitException(InitError::OutOfBoundary, "OutOfBoundary");
    auto &ps = state.particles[particle_id];
    const size_t per = cell == UnitCell::U ? 1 : 4;
    if (g[0] * g[1] * g[2] * per > ps.size()) throw InitException(InitError::TooBig, "TooBig");
    for (size_t x = 0; x < g[0]; ++x)
        for (size_t y = 0; y < g[1]; ++y)
            for (size_t z = 0; z < g[2]; ++z) {
                size_t c = x * g[1] * g[2] + y * g[2] + z;
                double fx = (double)x, fy = (double)y, fz = (double)z;
                if (cell == UnitCell::U) {
                    ps[c].position = {start[0] + fx * l, start[1] + fy * l, start[2] + fz * l};
                } else {
                    ps[4 * c].position = {start[0] + fx * l, start[1] + fy * l, start[2] + fz * l};
                    ps[4 * c + 1].position = {start[0] + fx * l, start[1] + (fy + 0.5) * l, start[2] + (fz + 0.5) * l};
                    ps[4 * c + 2].position = {start[0] + (fx + 0.5) * l, start[1] + fy * l, start[2] + (fz + 0.5) * l};
                    ps[4 * c + 3].position = {start[0] + (fx + 0.5) * l, start[1] + (fy + 0.5) * l, start[2] + fz * l};
                }
            }
}

void initialize_velocities_maxwell_boltzmann(State &state, double temperature, uint16_t particle_id, uint64_t seed)
{
    std::mt19937_64 rng(seed ? seed : std::random_device{}());
    std::normal_distribution<double> normal(0.0, 1.0);
    double t = temperature * 0.01;
    auto mass = ParticleDatabase::get_particle_mass(particle_id);
    if (!mass) throw Error(MD_ERR_INVALID_ARGUMENT, "No particle in DB");
    double sigma = std::sqrt(K_B * t / *mass);
    auto &ps = state.particles[particle_id];
    size_t half = ps.size() / 2;
    for (size_t i = 0; i < half; ++i) {
        double x = sigma * normal(rng), y = sigma * normal(rng), z = sigma * normal(rng);
        ps[i].velocity = {x, y, z};
        ps[i + half].velocity = {-x, -y, -z};
    }
}

}  // namespace initializer

}  // namespace moldyn
