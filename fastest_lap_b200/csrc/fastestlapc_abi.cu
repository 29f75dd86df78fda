// The reference's C API (include/fastestlapc.h <- /root/reference/src/main/c/fastestlapc.h:28-107) on top of the B200 path.
//
// Host C++ only: tables of named objects (fastestlapc.cpp:30-58), the XML loaders of the vehicle and track databases
// (create_vehicle_from_xml :103-142, create_track_from_xml :176-203; track format track_by_polynomial.hpp:78-153), the option
// parser of optimal_laptime (:1129-1440, subset), and the sequences of GPU solves behind optimal_laptime (:1442-1735) and
// gg_diagram (:1111-1127 -> Steady_state::gg_diagram, steady_state.hpp:696-880).  Every number that depends on the vehicle
// model comes out of the CUDA kernels through the C ABI of include/fastestlap_b200.h; there is no CPU model here.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fastestlap_b200.h"
#include "../../include/fastestlapc.h"

namespace flc {

using std::string;
using vec = std::vector<double>;
const double PI = 3.14159265358979323846, DEG = PI / 180.0, KMH = 1.0 / 3.6, G0 = 9.81;

struct error : std::runtime_error { using std::runtime_error::runtime_error; };

// ---------------------------------------------------------------------------------------------------------------------
// A small XML reader: elements, attributes, text; comments and the declaration are skipped.
struct Xml
{
    string tag, text;
    std::map<string, string> attr;
    std::vector<std::unique_ptr<Xml>> kids;

    const Xml* child(const string& t) const { for (auto& k : kids) if (k->tag == t) return k.get(); return nullptr; }
    const Xml* find(const string& path) const
    {
        const Xml* n = this;
        size_t a = 0;
        while (n && a <= path.size())
        {
            const size_t b = path.find('/', a);
            const string part = path.substr(a, b == string::npos ? string::npos : b - a);
            if (!part.empty()) n = n->child(part);
            if (b == string::npos) break;
            a = b + 1;
        }
        return n;
    }
};

static string trim(const string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

struct XmlParser
{
    const string& s;
    size_t i = 0;
    explicit XmlParser(const string& src) : s(src) {}
    void skip_misc()
    {
        for (;;)
        {
            while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
            if (s.compare(i, 4, "<!--") == 0) { const size_t e = s.find("-->", i); if (e == string::npos) throw error("XML: unterminated comment"); i = e + 3; }
            else if (s.compare(i, 2, "<?") == 0) { const size_t e = s.find("?>", i); if (e == string::npos) throw error("XML: unterminated declaration"); i = e + 2; }
            else if (s.compare(i, 2, "<!") == 0) { const size_t e = s.find('>', i); if (e == string::npos) throw error("XML: unterminated <!"); i = e + 1; }
            else return;
        }
    }
    std::unique_ptr<Xml> element()
    {
        skip_misc();
        if (i >= s.size() || s[i] != '<') throw error("XML: element expected");
        ++i;
        std::unique_ptr<Xml> n(new Xml);
        while (i < s.size() && !std::isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/') n->tag += s[i++];
        for (;;)
        {
            while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
            if (i >= s.size()) throw error("XML: unterminated tag <" + n->tag + ">");
            if (s[i] == '/') { if (s.compare(i, 2, "/>") != 0) throw error("XML: bad tag"); i += 2; return n; }
            if (s[i] == '>') { ++i; break; }
            string key;
            while (i < s.size() && s[i] != '=' && !std::isspace((unsigned char)s[i])) key += s[i++];
            while (i < s.size() && (std::isspace((unsigned char)s[i]) || s[i] == '=')) ++i;
            if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) throw error("XML: attribute value expected in <" + n->tag + ">");
            const char q = s[i++];
            const size_t e = s.find(q, i);
            if (e == string::npos) throw error("XML: unterminated attribute");
            n->attr[key] = s.substr(i, e - i);
            i = e + 1;
        }
        for (;;)
        {
            const size_t lt = s.find('<', i);
            if (lt == string::npos) throw error("XML: <" + n->tag + "> is not closed");
            n->text += s.substr(i, lt - i);
            i = lt;
            if (s.compare(i, 2, "</") == 0)
            {
                const size_t e = s.find('>', i);
                if (e == string::npos || trim(s.substr(i + 2, e - i - 2)) != n->tag) throw error("XML: mismatched </" + n->tag + ">");
                i = e + 1;
                n->text = trim(n->text);
                return n;
            }
            if (s.compare(i, 4, "<!--") == 0 || s.compare(i, 2, "<?") == 0) { skip_misc(); continue; }
            n->kids.push_back(element());
        }
    }
};

static std::unique_ptr<Xml> parse_xml(const string& text) { XmlParser p(text); return p.element(); }
static string read_file(const string& name)
{
    std::ifstream f(name);
    if (!f) throw error("cannot open \"" + name + "\"");
    std::stringstream ss; ss << f.rdbuf();
    return ss.str();
}
static vec numbers(const string& text)
{
    vec v;
    string t = text;
    for (char& c : t) if (c == ',' || c == '\n' || c == '\t') c = ' ';
    std::stringstream ss(t);
    double x;
    while (ss >> x) v.push_back(x);
    return v;
}
static bool truthy(const Xml* n) { if (!n) return false; string t = n->text; for (char& c : t) c = (char)std::tolower(c); return t == "true" || t == "1"; }

// ---------------------------------------------------------------------------------------------------------------------
// Vehicles: parameter paths in the order of the kernels' parameter vector (include/fastestlap_b200.h, FL_F1_* / FL_KART_*)
static const char* F1_TIRE[] = { "radius", "reference-load-1", "reference-load-2", "mu-x-max-1", "mu-x-max-2", "kappa-max-1", "kappa-max-2",
                                 "mu-y-max-1", "mu-y-max-2", "lambda-max-1", "lambda-max-2", "Qx", "Qy" };
static const char* F1_HEAD[] = { "chassis/mass", "chassis/inertia/Ixx", "chassis/inertia/Iyy", "chassis/inertia/Izz",
    "chassis/aerodynamics/rho", "chassis/aerodynamics/area", "chassis/aerodynamics/cd", "chassis/aerodynamics/cl",
    "chassis/com/x", "chassis/com/y", "chassis/com/z", "chassis/front_axle/x", "chassis/front_axle/z", "chassis/rear_axle/x", "chassis/rear_axle/z",
    "chassis/pressure_center/x", "chassis/pressure_center/y", "chassis/pressure_center/z",
    "chassis/brake_bias", "chassis/roll_balance_coefficient", "chassis/Fz_max_ref2",
    "front-axle/track", "front-axle/inertia", "front-axle/smooth_throttle_coeff", "front-axle/brakes/max_torque",
    "rear-axle/track", "rear-axle/inertia", "rear-axle/smooth_throttle_coeff", "rear-axle/differential_stiffness",
    "rear-axle/brakes/max_torque", "rear-axle/engine/maximum-power", "rear-axle/boost/maximum-power" };
static const char* KART_TIRE[] = { "radius", "radial-stiffness", "radial-damping", "nominal-vertical-load", "lambdaFz0", "Fz-max-ref2",
    "longitudinal/pure/pCx1", "longitudinal/pure/pDx1", "longitudinal/pure/pEx1", "longitudinal/pure/pKx1", "longitudinal/pure/pKx2", "longitudinal/pure/pKx3",
    "longitudinal/combined/rBx1", "longitudinal/combined/rCx1", "lateral/pure/pCy1", "lateral/pure/pDy1", "lateral/pure/pEy1", "lateral/pure/pKy1",
    "lateral/pure/pKy2", "lateral/pure/pKy4", "lateral/combined/rBy1", "lateral/combined/rCy1" };
static const char* KART_HEAD[] = { "chassis/mass", "chassis/inertia/Ixx", "chassis/inertia/Iyy", "chassis/inertia/Izz", "chassis/inertia/Ixz",
    "chassis/aerodynamics/rho", "chassis/aerodynamics/cd", "chassis/aerodynamics/cl", "chassis/aerodynamics/area",
    "chassis/com/x", "chassis/com/y", "chassis/com/z", "chassis/front_axle/x", "chassis/front_axle/z", "chassis/rear_axle/x", "chassis/rear_axle/z",
    "front-axle/track", "front-axle/stiffness/chassis", "front-axle/stiffness/antiroll", "front-axle/beta-steering/left", "front-axle/beta-steering/right",
    "rear-axle/track", "rear-axle/stiffness/chassis", "rear-axle/stiffness/antiroll", "rear-axle/inertia",
    "rear-axle/smooth_throttle_coeff", "rear-axle/brakes/max_torque", "rear-axle/engine/maximum-power" };
static const char* WIND[] = { "chassis/aerodynamics/wind_velocity/northward", "chassis/aerodynamics/wind_velocity/eastward" };

static std::vector<string> paths_of(int model)
{
    std::vector<string> p;
    if (model == FL_MODEL_F1_3DOF)
    {
        for (const char* h : F1_HEAD) p.push_back(h);
        for (const char* pre : { "front-tire/", "rear-tire/" }) for (const char* t : F1_TIRE) p.push_back(string(pre) + t);
    }
    else
    {
        for (const char* h : KART_HEAD) p.push_back(h);
        for (const char* pre : { "front-tire/", "rear-tire/" }) for (const char* t : KART_TIRE) p.push_back(string(pre) + t);
    }
    return p;
}

static const char* F1_INPUTS[] = { "front-axle.left-tire.kappa", "front-axle.right-tire.kappa", "rear-axle.left-tire.kappa", "rear-axle.right-tire.kappa",
    "chassis.velocity.x", "chassis.velocity.y", "chassis.omega.z", "chassis.force_z_fl_g", "chassis.force_z_fr_g", "chassis.force_z_rl_g",
    "chassis.force_z_rr_g", "time", "road.lateral-displacement", "road.track-heading-angle" };
static const char* F1_CONTROLS[] = { "front-axle.steering-angle", "rear-axle.boost", "chassis.throttle", "chassis.brake-bias" };
static const char* KART_INPUTS[] = { "rear-axle.omega", "chassis.velocity.x", "chassis.velocity.y", "chassis.omega.z", "chassis.position.z",
    "chassis.attitude.phi", "chassis.attitude.mu", "chassis.velocity.z", "chassis.omega.x", "chassis.omega.y", "time",
    "road.lateral-displacement", "road.track-heading-angle" };
static const char* KART_CONTROLS[] = { "front-axle.steering-angle", "rear-axle.throttle" };
static const char* OUTPUTS[] = { "laptime", "x", "y", "psi" };
static const char* KEY_NAME = "road.arclength";

struct VarParam { vec values; std::vector<int> idx; vec pts; std::vector<string> aliases; };      // aliases: one per value, or none

struct Vehicle
{
    int model = 0;
    vec params;
    double wind[2] = { 0.0, 0.0 };
    std::map<int, VarParam> variable;          // parameter slot -> values along the arclength
    string track;                              // vehicle_change_track (only recorded: the track is an argument of optimal_laptime)
    std::vector<std::pair<string, int>> declared;      // (alias, parameter slot): vehicle_declare_new_constant_parameter -- what <compute_sensitivity> differentiates by

    int slot(const string& path) const
    {
        const string key = path.compare(0, 8, "vehicle/") == 0 ? path.substr(8) : path;
        const auto p = paths_of(model);
        for (size_t k = 0; k < p.size(); ++k) if (p[k] == key) return (int)k;
        for (int k = 0; k < 2; ++k) if (key == WIND[k]) return -1 - k;
        throw error("Parameter \"" + path + "\" was not found");
    }
    void set(const string& path, double value) { const int k = slot(path); if (k < 0) wind[-1 - k] = value; else params[k] = value; }
    // Model_parameter::operator()(s), model_parameter.h:51-77
    vec params_at(double s) const
    {
        vec p = params;
        for (const auto& kv : variable)
        {
            const VarParam& v = kv.second;
            if (v.values.size() == 1 || v.pts.empty()) { p[kv.first] = v.values.front(); continue; }
            const size_t k = std::upper_bound(v.pts.begin(), v.pts.end(), s) - v.pts.begin();
            if (k == 0) p[kv.first] = v.values.front();
            else if (k == v.pts.size()) p[kv.first] = v.values.back();
            else
            {
                const double xi = (s - v.pts[k - 1]) / (v.pts[k] - v.pts[k - 1]);
                p[kv.first] = v.values[v.idx[k - 1]] + xi * (v.values[v.idx[k]] - v.values[v.idx[k - 1]]);
            }
        }
        return p;
    }
    string to_xml() const
    {
        // one leaf per parameter path (vehicle_save_as_xml); written as nested elements
        std::unique_ptr<Xml> root(new Xml);
        root->tag = "vehicle";
        root->attr["type"] = model == FL_MODEL_F1_3DOF ? "f1-3dof" : "kart-6dof";
        auto put = [&](const string& path, double value) {
            Xml* n = root.get();
            size_t a = 0;
            for (;;)
            {
                const size_t b = path.find('/', a);
                const string part = path.substr(a, b == string::npos ? string::npos : b - a);
                Xml* c = const_cast<Xml*>(n->child(part));
                if (!c) { n->kids.emplace_back(new Xml); c = n->kids.back().get(); c->tag = part; }
                n = c;
                if (b == string::npos) break;
                a = b + 1;
            }
            char buf[64]; std::snprintf(buf, sizeof buf, "%.17g", value); n->text = buf;
        };
        const auto p = paths_of(model);
        for (size_t k = 0; k < p.size(); ++k) put(p[k], params[k]);
        for (int k = 0; k < 2; ++k) if (wind[k] != 0.0) put(WIND[k], wind[k]);
        std::string out;
        std::function<void(const Xml&, int)> write = [&](const Xml& n, int depth) {
            out += string(4 * depth, ' ') + "<" + n.tag;
            for (auto& a : n.attr) out += " " + a.first + "=\"" + a.second + "\"";
            if (n.kids.empty()) { out += ">" + n.text + "</" + n.tag + ">\n"; return; }
            out += ">\n";
            for (auto& k : n.kids) write(*k, depth + 1);
            out += string(4 * depth, ' ') + "</" + n.tag + ">\n";
        };
        write(*root, 0);
        return out;
    }
};

// value of `path` below <vehicle>: three-vectors are written as <x>,<y>,<z> children (F1 files) or as one text node (kart files)
static bool lookup(const Xml& root, const string& path, double& value)
{
    const Xml* n = root.find(path);
    if (n && !n->text.empty()) { const vec v = numbers(n->text); if (v.size() == 1) { value = v[0]; return true; } }
    const size_t cut = path.rfind('/');
    if (cut == string::npos) return false;
    const Xml* parent = root.find(path.substr(0, cut));
    const string leaf = path.substr(cut + 1);
    const int c = leaf == "x" ? 0 : leaf == "y" ? 1 : leaf == "z" ? 2 : -1;
    if (parent && c >= 0) { const vec v = numbers(parent->text); if (v.size() == 3) { value = v[c]; return true; } }
    return false;
}

static Vehicle vehicle_from_xml(const string& text)
{
    const auto root = parse_xml(text);
    Vehicle v;
    const string type = root->attr.count("type") ? root->attr.at("type") : "";
    if (type == "f1-3dof") v.model = FL_MODEL_F1_3DOF;
    else if (type == "kart-6dof") v.model = FL_MODEL_KART_6DOF;
    else throw error("vehicle type \"" + type + "\" not recognized");
    for (const string& p : paths_of(v.model))
    {
        double x;
        if (!lookup(*root, p, x)) throw error("vehicle parameter \"" + p + "\" is missing");
        v.params.push_back(x);
    }
    for (int k = 0; k < 2; ++k) { double x; if (lookup(*root, WIND[k], x)) v.wind[k] = x; }
    for (const char* axle : { "rear-axle", "front-axle" })
        if (root->find(string(axle) + "/engine/rpm-data") || root->find(string(axle) + "/engine/power-data"))
            throw error(string("engine power curves (\"") + axle + "/engine/rpm-data\") are not supported: only engine/maximum-power");
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// Tracks: the discrete format of the database (track_by_polynomial.hpp:78-153): piecewise-linear in the arclength
struct Track
{
    bool closed = true, elevation = false;
    double length = 0.0;
    vec s;                                       // as in the file (closed: without the repeated first point)
    std::map<string, vec> a;                     // x y z yaw pitch roll yaw_dot dpitch droll nl nr

    double at(const string& key, double sq) const
    {
        const vec& v = a.at(key);
        const size_t n = s.size();
        if (sq <= s[0]) return v[0];
        const size_t k = std::upper_bound(s.begin(), s.end(), sq) - s.begin();
        if (k < n) { const double xi = (sq - s[k - 1]) / (s[k] - s[k - 1]); return v[k - 1] + xi * (v[k] - v[k - 1]); }
        if (!closed) return v[n - 1];
        if (sq >= length) return v[0];
        const double xi = (sq - s[n - 1]) / (length - s[n - 1]);          // the closing element: back to the first sample
        return v[n - 1] + xi * (v[0] - v[n - 1]);
    }
    bool has_elevation() const
    {
        for (const char* k : { "pitch", "roll", "dpitch", "droll" }) for (double x : a.at(k)) if (x != 0.0) return true;
        for (double x : a.at("z")) if (std::fabs(x) >= 1.0e-12) return true;
        return false;
    }
};

static Track track_from_xml(const string& text)
{
    const auto root = parse_xml(text);
    auto att = [&](const char* k) { return root->attr.count(k) ? root->attr.at(k) : string(); };
    if (att("format") != "discrete") throw error("Track should be of type \"discrete\"");
    if (att("type") != "open" && att("type") != "closed") throw error("Incorrect track type, should be \"open\" or \"closed\"");
    if (att("dimensions") != "2" && att("dimensions") != "3") throw error("Track XML file must have a dimensions attribute equal to 2 or 3");
    Track t;
    t.closed = att("type") == "closed";
    const bool elev = att("dimensions") == "3";
    const Xml* data = root->child("data");
    if (!data) throw error("track file has no <data>");
    const size_t n = (size_t)std::stol(data->attr.at("number_of_points"));
    auto arr = [&](const char* path, bool required) -> vec {
        const Xml* e = data->find(path);
        if (!e) { if (required) throw error(string("track file has no <") + path + ">"); return vec(n, 0.0); }
        vec v = numbers(e->text);
        if (v.size() != n) throw error("track arrays must have number_of_points entries");
        return v;
    };
    t.s = arr("arclength", true);
    t.a["x"] = arr("centerline/x", true); t.a["y"] = arr("centerline/y", true); t.a["z"] = arr("centerline/z", false);
    t.a["yaw"] = arr("yaw", true); t.a["yaw_dot"] = arr("yaw_dot", true); t.a["nl"] = arr("nl", true); t.a["nr"] = arr("nr", true);
    t.a["pitch"] = arr("pitch", false); t.a["roll"] = arr("roll", false); t.a["dpitch"] = arr("dpitch_dot", false); t.a["droll"] = arr("droll_dot", false);
    if (!elev) for (const char* k : { "z", "pitch", "roll", "dpitch", "droll" }) t.a[k] = vec(n, 0.0);
    const Xml* len = root->find("header/track_length");
    if (!len) throw error("track file has no header/track_length");
    t.length = numbers(len->text).at(0);
    // kerbs (track_by_polynomial.hpp:93-108): a sample inside a used kerb gets nl / nr widened by the width of the first such kerb
    if (const Xml* kerbs = root->child("kerbs"))
        for (const char* side : { "left", "right" })
            if (const Xml* sd = kerbs->child(side))
            {
                vec& w = t.a[string(side) == "left" ? "nl" : "nr"];
                std::vector<char> done(n, 0);
                for (auto& kerb : sd->kids)
                {
                    string used = kerb->attr.count("used") ? kerb->attr.at("used") : "";
                    for (char& c : used) c = (char)std::tolower(c);
                    if (trim(used) != "1" && trim(used) != "true") continue;
                    const double s0 = numbers(kerb->child("arclength_start")->text).at(0), s1 = numbers(kerb->child("arclength_finish")->text).at(0);
                    const double wd = numbers(kerb->child("width")->text).at(0);
                    for (size_t i = 0; i < n; ++i)
                        if (!done[i] && t.s[i] > s0 - 1.0e-12 && t.s[i] < s1 + 1.0e-12) { w[i] += wd; done[i] = 1; }
                }
            }
    t.elevation = elev;
    return t;
}

// ---------------------------------------------------------------------------------------------------------------------
// tables (fastestlapc.cpp:30-58)
static std::map<string, Vehicle> vehicles;
static std::map<string, Track> tracks;
static std::map<string, double> scalars;
static std::map<string, vec> vectors;
static string last_error;
static int print_level = 0;

static void check_new(const string& name)
{
    if (vehicles.count(name)) throw error(string("Vehicle of type ") + (vehicles[name].model == FL_MODEL_F1_3DOF ? "f1-3dof" : "kart-6dof") + " with name \"" + name + "\" already exists");
    if (tracks.count(name)) throw error("Track with name \"" + name + "\" already exists");
    if (scalars.count(name)) throw error("Scalar with name \"" + name + "\" already exists");
    if (vectors.count(name)) throw error("Vector with name \"" + name + "\" already exists");
}
static Vehicle& vehicle(const string& name) { auto it = vehicles.find(name); if (it == vehicles.end()) throw error("Vehicle \"" + name + "\" does not exist"); return it->second; }
static Track& track(const string& name) { auto it = tracks.find(name); if (it == tracks.end()) throw error("Track \"" + name + "\" does not exist"); return it->second; }
static string fl_error() { const char* e = fl_last_error(); return e ? e : ""; }

// ---------------------------------------------------------------------------------------------------------------------
// GPU solves
struct SteadyResult { std::vector<vec> x; std::vector<int> status; };

// A batch of steady-state NLPs (fl_steady_create + the interior-point solver); bounds: empty (model defaults), one set, or one per problem
static SteadyResult steady_batch(int model, const std::vector<vec>& params, const std::vector<vec>& spec, const std::vector<vec>& x0,
                                 const vec& xl = vec(), const vec& xu = vec(), const vec& gl = vec(), const vec& gu = vec())
{
    const int B = (int)spec.size();
    const int nvp = fl_n_vehicle_params(model);
    vec P((size_t)B * nvp), S((size_t)B * 7);
    for (int b = 0; b < B; ++b)
    {
        const vec& pb = params.size() == 1 ? params[0] : params[b];
        std::copy(pb.begin(), pb.end(), P.begin() + (size_t)b * nvp);
        std::copy(spec[b].begin(), spec[b].end(), S.begin() + (size_t)b * 7);
    }
    fl_steady_desc d;
    d.model = model; d.batch = B; d.vehicle_params = P.data(); d.spec = S.data(); d.device = 0;
    fl_nlp* nlp = fl_steady_create(&d);
    if (!nlp) throw error("steady-state problem: " + fl_error());
    int n = 0, m = 0;
    fl_nlp_sizes(nlp, &n, &m, nullptr, nullptr);
    fl_ipm_opts o;
    fl_ipm_default_options(&o);
    o.tol = 1.0e-8; o.constr_viol_tol = 1.0e-8; o.acceptable_tol = 1.0e-6; o.max_iter = 300;          // steady_state.hpp:84-88
    const bool per_lap = (int)xl.size() == B * n && B > 1;
    fl_ipm* s = xl.empty() ? fl_ipm_create(nlp, &o, nullptr, nullptr, nullptr, nullptr) : fl_ipm_create_ex(nlp, &o, xl.data(), xu.data(), gl.data(), gu.data(), per_lap ? 1 : 0);
    if (!s) { fl_nlp_destroy(nlp); throw error("steady-state solver: " + fl_error()); }
    vec X0((size_t)B * n), X((size_t)B * n);
    for (int b = 0; b < B; ++b) std::copy(x0[b].begin(), x0[b].begin() + n, X0.begin() + (size_t)b * n);
    std::vector<int> status(B), iters(B);
    const int ok = fl_ipm_solve(s, X0.data(), 0) && fl_ipm_results(s, X.data(), nullptr, nullptr, nullptr, nullptr, status.data(), iters.data(), nullptr);
    const string msg = ok ? "" : fl_error();
    fl_ipm_destroy(s); fl_nlp_destroy(nlp);
    if (!ok) throw error("steady-state solve: " + msg);
    SteadyResult r;
    r.status = status;
    for (int b = 0; b < B; ++b) r.x.emplace_back(X.begin() + (size_t)b * n, X.begin() + (size_t)(b + 1) * n);
    return r;
}

static vec steady_guess(int model) { return model == FL_MODEL_F1_3DOF ? vec{ 0, 0, 0, 0, 0, -0.25, -0.25, -0.25, -0.25, 0, 0, 0, 0 } : vec{ 0, 0.02, 0, 0, 0, 0, 0, 0 }; }

// the straight-line steady state at `speed` as one node of the lap NLP (fastestlapc.cpp:1508-1545)
static vec start_node(const Vehicle& veh, double speed, bool derivative_form)
{
    const SteadyResult r = steady_batch(veh.model, { veh.params }, { vec{ speed, 0, 0, 0, 0, 0, 0 } }, { steady_guess(veh.model) });
    if (r.status[0] < 1) throw error("[ERROR] steady-state computation did not converge");
    const vec& x = r.x[0];
    vec node;
    if (veh.model == FL_MODEL_F1_3DOF)
    {
        const double ang = 10.0 * DEG;
        node = { x[0], x[1], x[2], x[3], speed * std::cos(x[4] * ang), -speed * std::sin(x[4] * ang), 0.0, x[5], x[6], x[7], x[8], 0.0, x[4] * ang, x[9] * ang, x[10] };
    }
    else
    {
        const double psi = x[4];
        node = { (x[0] + 1.0) * speed / 0.139, speed * std::cos(psi), -speed * std::sin(psi), 0.0, x[1], x[2], x[3], 0.0, 0.0, 0.0, 0.0, psi, x[5], 0.0 };
    }
    if (derivative_form) { node.push_back(0.0); node.push_back(0.0); }
    return node;
}

struct LapOptions
{
    double speed = 50.0, sigma = 0.5, diss[2] = { 0.0, 0.0 };
    bool have_diss = false;
    string prefix;
    int print_level = 0;
    // one quantity that couples the whole lap, carried node-wise (include/fastestlap_b200.h, aug_*): the brake bias as a constant /
    // on a hypermesh (fastestlapc.cpp:1276-1283) or one restricted integral quantity (:1298-1306)
    std::vector<string> variables;             // <output_variables><variables>: names asked for (fastestlapc.cpp:1215-1233)
    bool closed = true, warm_start = false, save_warm_start = false;
    bool sensitivity = false;                  // <compute_sensitivity> (fastestlapc.cpp:1254, 1525)
    vec q_start, u_start;                      // open simulations: <initial_condition><q from_table/><u from_table/> (fastestlapc.cpp:1237-1249)
    int aug_kind = FL_AUG_NONE;
    vec hypermesh;
    string integral_name;
    double integral_bounds[2] = { 0.0, 0.0 };
};

static LapOptions lap_options(const string& text, int model)
{
    LapOptions o;
    // default dissipations (fastestlapc.cpp:1412-1440)
    if (model == FL_MODEL_F1_3DOF) { o.diss[0] = 50.0; o.diss[1] = 20.0 * 8.0e-4; } else { o.diss[0] = 1.0e-2; o.diss[1] = 200.0 * 200.0 * 1.0e-10; }
    if (trim(text).empty()) return o;
    const auto root = parse_xml(text);
    o.sensitivity = truthy(root->child("compute_sensitivity"));
    int n_coupling = 0;
    if (const Xml* ic = root->child("integral_constraints"))
        for (auto& q : ic->kids)
        {
            if (!q->child("lower_bound") || !q->child("upper_bound")) throw error("[ERROR] optimal_laptime -> integral constraint \"" + q->tag + "\" needs <lower_bound> and <upper_bound>");
            o.aug_kind = FL_AUG_INTEGRAL; o.integral_name = q->tag;
            o.integral_bounds[0] = numbers(q->child("lower_bound")->text).at(0); o.integral_bounds[1] = numbers(q->child("upper_bound")->text).at(0);
            ++n_coupling;
        }
    if (const Xml* n = root->child("closed_simulation")) o.closed = truthy(n);
    o.warm_start = truthy(root->child("warm_start"));
    o.save_warm_start = truthy(root->child("save_warm_start"));
    if (!o.closed)
    {
        const Xml* ic = root->child("initial_condition");
        if (!ic || !ic->child("q") || !ic->child("u")) throw error("For open simulations, the initial condition must be providedin 'options/initial_condition'");      // (the reference's message)
        for (int k = 0; k < 2; ++k)
        {
            const Xml* n = ic->child(k == 0 ? "q" : "u");
            const string from = n->attr.count("from_table") ? n->attr.at("from_table") : "";
            if (!vectors.count(from)) throw error("[ERROR] optimal_laptime -> initial condition: vector \"" + from + "\" does not exist");
            (k == 0 ? o.q_start : o.u_start) = vectors.at(from);
        }
    }
    if (const Xml* n = root->child("initial_speed")) o.speed = numbers(n->text).at(0);
    if (const Xml* n = root->child("steady_state_speed")) o.speed = numbers(n->text).at(0);
    if (const Xml* n = root->child("sigma")) o.sigma = numbers(n->text).at(0);
    if (const Xml* n = root->child("print_level")) o.print_level = (int)numbers(n->text).at(0);
    if (const Xml* ov = root->child("output_variables")) if (const Xml* p = ov->child("prefix")) o.prefix = trim(p->text);
    if (const Xml* ov = root->child("output_variables")) if (const Xml* vs = ov->child("variables")) for (auto& v : vs->kids) o.variables.push_back(v->tag);
    if (const Xml* cv = root->child("control_variables"))
    {
        vec d;
        for (auto& c : cv->kids)
        {
            const string type = c->attr.count("optimal_control_type") ? c->attr.at("optimal_control_type") : "";
            if (type == "constant" || type == "hypermesh")
            {
                if (model != FL_MODEL_F1_3DOF || c->tag != "chassis.brake-bias")
                    throw error("[ERROR] optimal_laptime -> control \"" + c->tag + "\": only the F1's chassis.brake-bias can be optimised as a constant / on a hypermesh");
                if (type == "hypermesh" && !c->child("hypermesh")) throw error("[ERROR] optimal_laptime -> control variable \"" + c->tag + "\" needs its <hypermesh>");
                o.aug_kind = FL_AUG_BRAKE_BIAS;
                o.hypermesh = type == "constant" ? vec{ 0.0 } : numbers(c->child("hypermesh")->text);
                ++n_coupling;
            }
            else if (type == "full-mesh") { if (const Xml* q = c->child("dissipation")) d.push_back(numbers(q->text).at(0)); }
            else if (type != "dont optimize") throw error("[ERROR] Optimal control type \"" + type + "\" not recognized");
        }
        if (d.size() == 2) { o.diss[0] = d[0]; o.diss[1] = d[1]; }
    }
    if (n_coupling > 1) throw error("[ERROR] optimal_laptime -> one hypermesh / constant control or one integral constraint per optimal_laptime call");
    return o;
}

static void put_vector(const string& name, const vec& v) { vectors[name] = v; }

// <save_warm_start> keeps the last solution per vehicle type, <warm_start> restarts from it with its multipliers on ITS mesh and with the
// vehicle as it is now (fastestlapc.cpp:60-71, 1561-1568, 1712-1713)
struct WarmStart { bool valid = false; string track; vec s, x, lam, zl, zu; };
static WarmStart warm_starts[2];

static void do_optimal_laptime(const string& vname, const string& tname, const vec& s_in, const string& options)
{
    Vehicle& veh = vehicle(vname);
    Track& trk = track(tname);
    const LapOptions o = lap_options(options, veh.model);
    const bool f1 = veh.model == FL_MODEL_F1_3DOF;
    WarmStart& ws = warm_starts[f1 ? 0 : 1];
    if (o.warm_start)
    {
        if (!ws.valid) throw error("[ERROR] optimal_laptime -> <warm_start> without a simulation saved by <save_warm_start>");
        if (ws.track != tname) throw error("[ERROR] optimal_laptime -> the warm start was saved for another track");
        if (!o.closed || o.aug_kind != FL_AUG_NONE) throw error("[ERROR] optimal_laptime -> <warm_start> of open simulations / with hypermesh controls or integral constraints is not supported");
    }
    if (o.save_warm_start && (!o.closed || o.aug_kind != FL_AUG_NONE))
        throw error("[ERROR] optimal_laptime -> <save_warm_start> of open simulations / with hypermesh controls or integral constraints is not supported");
    if (!o.closed && o.aug_kind != FL_AUG_NONE) throw error("[ERROR] optimal_laptime -> hypermesh controls and integral constraints need a closed simulation");
    if (o.sensitivity && o.aug_kind != FL_AUG_NONE)
        throw error("[ERROR] optimal_laptime -> <compute_sensitivity> is not offered together with hypermesh controls or integral constraints");
    // what the analysis differentiates by: the declared constants, and every value of the parameters that vary along the lap (one alias each)
    struct SensPar { string alias; int slot; int value; };          // value < 0: a constant
    std::vector<SensPar> sens;
    for (auto& dpar : veh.declared) sens.push_back({ dpar.first, dpar.second, -1 });
    for (auto& kv : veh.variable) for (size_t k = 0; k < kv.second.aliases.size(); ++k) sens.push_back({ kv.second.aliases[k], kv.first, (int)k });
    const vec s = o.warm_start ? ws.s : s_in;
    const int np = (int)s.size();
    if (np < 2) throw error("[ERROR] optimal_laptime -> provide at least two values of arclength");
    const int form = f1 ? FL_FORM_DIRECT : FL_FORM_DERIVATIVE;
    const bool elevation = trk.has_elevation();
    if (elevation && !f1) throw error("lot2016kart does not support tracks with elevation");
    if (o.closed && s.back() >= trk.length) throw error("closed meshes must not repeat the first point");
    if (!o.closed && ((int)o.q_start.size() != (f1 ? 14 : 13) || (int)o.u_start.size() != (f1 ? 4 : 2)))
        throw error("[ERROR] optimal_laptime -> the initial condition must have " + std::to_string(f1 ? 14 : 13) + " states and " + std::to_string(f1 ? 4 : 2) + " controls");
    vec nodes((size_t)np * 6), wl(np), wr(np);
    const char* keys[6] = { "yaw", "pitch", "roll", "yaw_dot", "dpitch", "droll" };
    for (int i = 0; i < np; ++i)
    {
        for (int c = 0; c < 6; ++c) nodes[(size_t)i * 6 + c] = trk.at(keys[c], s[i]);
        wl[i] = trk.at("nl", s[i]); wr[i] = trk.at("nr", s[i]);
    }
    const int nvp = fl_n_vehicle_params(veh.model);
    vec node_params;
    if (!veh.variable.empty())
        for (int i = 0; i < np; ++i) { const vec p = veh.params_at(s[i]); node_params.insert(node_params.end(), p.begin(), p.end()); }
    Vehicle v0 = veh;
    if (!veh.variable.empty()) v0.params = veh.params_at(s[0]);
    const int n_in = f1 ? 14 : 13, n_ctrl = f1 ? 4 : 2;
    vec ctrl_default = f1 ? vec{ 0.0, 0.0, 0.0, v0.params[18] } : vec{ 0.0, 0.0 };
    fl_problem_desc d;
    std::memset(&d, 0, sizeof d);
    const double du0[2] = { 0.0, 0.0 };
    d.model = veh.model; d.form = form; d.closed = o.closed ? 1 : 0; d.elevation = elevation ? 1 : 0; d.kart_direct_torque = 0; d.n_points = np;
    if (!o.closed) { d.q0 = o.q_start.data(); d.u0 = o.u_start.data(); d.du0 = du0; }
    d.s = s.data(); d.track_length = trk.length; d.track_nodes = nodes.data(); d.wl = wl.data(); d.wr = wr.data();
    d.batch = 1; d.vehicle_params = v0.params.data(); d.sigma = o.sigma; d.dissipation[0] = o.diss[0]; d.dissipation[1] = o.diss[1];
    d.ctrl_default = ctrl_default.data(); d.device = 0;
    d.node_params = node_params.empty() ? nullptr : node_params.data();
    d.wind[0] = veh.wind[0]; d.wind[1] = veh.wind[1];
    (void)nvp;
    vec jump(np, 0.0);
    static const char* IQ[5] = { "engine-energy", "tire-fl-energy", "tire-fr-energy", "tire-rl-energy", "tire-rr-energy" };
    if (o.aug_kind == FL_AUG_BRAKE_BIAS)
    {
        // stretch of a node = the last hypermesh point <= s (control_array_at_s, optimal_laptime_control_variables.h:190-222)
        const vec& sh = o.hypermesh;
        if (sh.empty() || std::fabs(sh[0]) > 1.0e-12 || sh.back() >= trk.length) throw error("[ERROR] optimal_laptime -> the hypermesh starts at s = 0 and ends before the track length");
        for (size_t k = 1; k < sh.size(); ++k) if (!(sh[k] > sh[k - 1])) throw error("[ERROR] optimal_laptime -> the hypermesh must increase strictly");
        std::vector<int> stretch(np);
        for (int i = 0; i < np; ++i) stretch[i] = (int)(std::upper_bound(sh.begin(), sh.end(), s[i]) - sh.begin()) - 1;
        for (int i = 0; i < np; ++i) jump[i] = stretch[i] != stretch[(i + np - 1) % np] ? 1.0 : 0.0;
        d.aug_kind = FL_AUG_BRAKE_BIAS; d.aug_jump = jump.data(); d.aug_bounds[0] = 0.0; d.aug_bounds[1] = 1.0;
    }
    else if (o.aug_kind == FL_AUG_INTEGRAL)
    {
        int which = -1;
        for (int k = 0; k < 5; ++k) if (o.integral_name == IQ[k]) which = k;
        if (!f1 || which < 0) throw error("[ERROR] optimal_laptime -> integral quantity \"" + o.integral_name + "\" is not offered (engine-energy, tire-fl/-fr/-rl/-rr-energy of the F1)");
        d.aug_kind = FL_AUG_INTEGRAL; d.aug_weights[which] = 1.0; d.aug_bounds[0] = o.integral_bounds[0]; d.aug_bounds[1] = o.integral_bounds[1];
    }
    fl_nlp* nlp = fl_nlp_create(&d);
    if (!nlp) throw error("[ERROR] optimal_laptime -> " + fl_error());
    int n = 0, m = 0;
    fl_nlp_sizes(nlp, &n, &m, nullptr, nullptr);
    const int nvn = o.closed ? np : np - 1;          // nodes that carry variables
    const int NV = n / nvn;
    vec X(n), Ylam(m), Zl(n), Zu(n), sigx, sigs;
    double dc_last = 0.0;
    std::vector<int> status(1, 0), iters(1, 0);
    double obj = 0.0;
    string msg;
    // the requested starting speed first; without a restoration phase a lap that does not converge from there is tried again with
    // a larger initial barrier parameter from the same and from faster steady states (as fastest_lap_b200/api.py does)
    const double factors[4] = { 1.0, 1.0, 2.0, 3.0 };
    for (int attempt = 0; attempt < 4 && status[0] < 1; ++attempt)
    {
        fl_ipm_opts io;
        fl_ipm_default_options(&io);
        if (attempt > 0) io.mu_init = 1.0;
        vec node = start_node(v0, factors[attempt] * o.speed * KMH, form == FL_FORM_DERIVATIVE);
        if (!o.closed)
        {
            // the given first node replicated over the free nodes (fastestlapc.cpp:1546-1557); control rates 0
            node.clear();
            for (int j = 0; j < n_in; ++j) if (j != n_in - 3) node.push_back(o.q_start[j]);
            node.push_back(o.u_start[0]); node.push_back(o.u_start[f1 ? 2 : 1]);
            if (form == FL_FORM_DERIVATIVE) { node.push_back(0.0); node.push_back(0.0); }
        }
        if (o.aug_kind != FL_AUG_NONE)      // (a, q) after the inputs: the vehicle's brake bias / the middle of the integral's bounds; no jump
            node.insert(node.begin() + (n_in - 1), { o.aug_kind == FL_AUG_BRAKE_BIAS ? std::min(std::max(v0.params[18], 0.0), 1.0) : 0.5 * (o.integral_bounds[0] + o.integral_bounds[1]), 0.0 });
        vec x0(n);
        for (int i = 0; i < nvn; ++i) std::copy(node.begin(), node.begin() + NV, x0.begin() + (size_t)i * NV);
        fl_ipm* ipm = fl_ipm_create(nlp, &io, nullptr, nullptr, nullptr, nullptr);
        if (!ipm) { msg = fl_error(); break; }
        int ok;
        if (o.warm_start) ok = fl_ipm_solve_warm(ipm, ws.x.data(), ws.lam.data(), ws.zl.data(), ws.zu.data(), 1.0e-3, 1.0e-4, o.print_level >= 5 ? 1 : 0);
        else ok = fl_ipm_solve(ipm, x0.data(), o.print_level >= 5 ? 1 : 0);
        if (!ok || !fl_ipm_results(ipm, X.data(), Ylam.data(), Zl.data(), Zu.data(), &obj, status.data(), iters.data(), nullptr)) msg = fl_error();
        if (msg.empty() && status[0] >= 1 && o.sensitivity && !sens.empty())
        {
            // barrier terms of the solution, with the solver's own bounds and slacks: the KKT matrix the sensitivities are solved with
            sigx.resize(n); sigs.resize((size_t)(o.closed ? np : np - 1) * 6);
            if (!fl_ipm_sigma(ipm, sigx.data(), sigs.data(), &dc_last)) msg = fl_error();
        }
        fl_ipm_destroy(ipm);
        if (!msg.empty() || o.warm_start || !o.closed) break;          // (the retries are for closed laps from the steady-state start)
    }
    // output channels of the model along the lap, at the solution (fastestlapc.cpp:1643-1690)
    std::map<string, vec> channels;
    if (msg.empty() && status[0] >= 1 && !o.variables.empty())
    {
        const int n_out = fl_nlp_n_outputs(nlp);
        std::vector<int> wanted;
        for (int k = 0; k < n_out; ++k) if (std::find(o.variables.begin(), o.variables.end(), string(fl_nlp_output_name(nlp, k))) != o.variables.end()) wanted.push_back(k);
        if (!wanted.empty())
        {
            vec out((size_t)n_out * np);
            if (!fl_eval_outputs(nlp, X.data(), out.data())) msg = fl_error();
            else for (int k : wanted) channels[fl_nlp_output_name(nlp, k)] = vec(out.begin() + (size_t)k * np, out.begin() + (size_t)(k + 1) * np);
        }
    }
    // Parameter sensitivities (Optimal_laptime with check_optimality, optimal_laptime.hpp:564-588): dx/dp = -K^-1 d/dp (grad f + J^T lambda, g)
    // at fixed (x, lambda), K = the KKT matrix of the solution factorised by the device kernels; the right-hand side is a central
    // difference of the GPU callbacks over the parameter (relative step 1e-6: the callbacks are closed-form in the parameters).
    std::vector<vec> dx_dp;
    if (msg.empty() && status[0] >= 1 && !sigx.empty())
    {
        int nnzj = 0, nnzh = 0;
        fl_nlp_sizes(nlp, nullptr, nullptr, &nnzj, &nnzh);
        std::vector<int> jr(nnzj), jc(nnzj), hr(nnzh), hc(nnzh);
        fl_ipm* ipm = fl_ipm_create(nlp, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (!ipm || !fl_nlp_structure(nlp, jr.data(), jc.data(), hr.data(), hc.data())) msg = fl_error();
        for (size_t q = 0; msg.empty() && q < sens.size(); ++q)
        {
            const int slot = sens[q].slot;
            const double base_value = sens[q].value < 0 ? v0.params[slot] : veh.variable.at(slot).values[sens[q].value];
            const double step = 1.0e-6 * std::max(std::fabs(base_value), 1.0e-3);
            vec gl[2], gv[2];
            for (int side = 0; side < 2 && msg.empty(); ++side)
            {
                Vehicle v2 = veh;
                if (sens[q].value < 0) v2.params[slot] += side == 0 ? step : -step;
                else v2.variable[slot].values[sens[q].value] += side == 0 ? step : -step;
                vec np2;
                if (!v2.variable.empty())
                    for (int i = 0; i < np; ++i) { const vec pi = v2.params_at(s[i]); np2.insert(np2.end(), pi.begin(), pi.end()); }
                vec pp = v2.variable.empty() ? v2.params : v2.params_at(s[0]);
                vec cd = ctrl_default;
                if (f1) cd[3] = pp[18];
                fl_problem_desc d2 = d;
                d2.vehicle_params = pp.data(); d2.ctrl_default = cd.data();
                d2.node_params = np2.empty() ? nullptr : np2.data();
                fl_nlp* h2 = fl_nlp_create(&d2);
                if (!h2) { msg = fl_error(); break; }
                vec g(m), grad(n), jac(nnzj);
                if (!fl_eval_all(h2, X.data(), Ylam.data(), 1.0, nullptr, g.data(), grad.data(), jac.data(), nullptr)) msg = fl_error();
                fl_nlp_destroy(h2);
                for (int k = 0; k < nnzj; ++k) grad[jc[k]] += jac[k] * Ylam[jr[k]];
                gl[side] = grad; gv[side] = g;
            }
            if (!msg.empty()) break;
            vec r1(n), r2(m), dx(n), dy(m);
            for (int i = 0; i < n; ++i) r1[i] = -(gl[0][i] - gl[1][i]) / (2.0 * step);
            for (int i = 0; i < m; ++i) r2[i] = -(gv[0][i] - gv[1][i]) / (2.0 * step);
            const double dw0 = 0.0;
            if (!fl_kkt_factor_solve(ipm, X.data(), Ylam.data(), sigx.data(), sigs.data(), &dw0, &dc_last, r1.data(), r2.data(), 3, 0, dx.data(), dy.data(), nullptr, nullptr)) msg = fl_error();
            dx_dp.push_back(dx);
        }
        if (ipm) fl_ipm_destroy(ipm);
    }
    fl_nlp_destroy(nlp);
    if (!msg.empty()) throw error("[ERROR] optimal_laptime -> " + msg);
    if (status[0] < 1) throw error("[ERROR] optimal_laptime -> the optimization did not converge (status " + std::to_string(status[0]) + ")");
    if (o.save_warm_start) { ws.valid = true; ws.track = tname; ws.s = s; ws.x = X; ws.lam = Ylam; ws.zl = Zl; ws.zu = Zu; }
    double t_start = 0.0;
    if (!o.closed)
    {
        // the fixed first node goes back in front of the free ones
        vec first;
        for (int j = 0; j < n_in; ++j) if (j != n_in - 3) first.push_back(o.q_start[j]);
        first.push_back(o.u_start[0]); first.push_back(o.u_start[f1 ? 2 : 1]);
        first.resize(NV, 0.0);
        X.insert(X.begin(), first.begin(), first.end());
        t_start = o.q_start[n_in - 3];
    }

    // the augmented variables leave the node: what follows reads the reference's layout
    vec aug_values;
    int NVo = NV;
    if (o.aug_kind != FL_AUG_NONE)
    {
        NVo = NV - 2;
        vec Y((size_t)np * NVo);
        aug_values.resize(np);
        for (int i = 0; i < np; ++i)
        {
            const double* x = X.data() + (size_t)i * NV;
            aug_values[i] = x[n_in - 1];
            std::copy(x, x + (n_in - 1), Y.begin() + (size_t)i * NVo);
            std::copy(x + (n_in + 1), x + NV, Y.begin() + (size_t)i * NVo + (n_in - 1));
        }
        X = Y;
    }
    // ---- outputs (optimal_laptime.hpp:598-631): time by the trapezoid rule of dt/ds, position = centreline + n * normal, psi = alpha + theta
    const int iu = f1 ? 4 : 1, iv = f1 ? 5 : 2, n_col = n_in - 3, a_col = n_in - 2;
    vec dtds(np), time(np, 0.0), xs(np), ys(np), psi(np);
    for (int i = 0; i < np; ++i)
    {
        const double* x = X.data() + (size_t)i * NVo;
        const double th = nodes[(size_t)i * 6], mu = nodes[(size_t)i * 6 + 1], ph = nodes[(size_t)i * 6 + 2];
        const double kz = -std::sin(ph) * nodes[(size_t)i * 6 + 4] + std::cos(mu) * std::cos(ph) * nodes[(size_t)i * 6 + 3];
        dtds[i] = (1.0 - x[n_col] * kz) / (x[iu] * std::cos(x[a_col]) - x[iv] * std::sin(x[a_col]));
        const double nx = std::cos(th) * std::sin(mu) * std::sin(ph) - std::sin(th) * std::cos(ph);
        const double ny = std::sin(th) * std::sin(mu) * std::sin(ph) + std::cos(th) * std::cos(ph);
        xs[i] = trk.at("x", s[i]) + x[n_col] * nx; ys[i] = trk.at("y", s[i]) + x[n_col] * ny; psi[i] = th + x[a_col];
    }
    double laptime = t_start;
    time[0] = t_start;
    for (int i = 0; i < (o.closed ? np : np - 1); ++i)
    {
        const int j = (i + 1) % np;
        const double h = (i + 1 < np) ? s[i + 1] - s[i] : trk.length - s[np - 1];
        laptime += h * 0.5 * (dtds[i] + dtds[j]);
        if (i + 1 < np) time[i + 1] = laptime;
    }
    const string pre = o.prefix;
    for (auto* tab : { (void*)0 }) (void)tab;
    auto overwrite = [&](const string& name) { vectors.erase(name); scalars.erase(name); };
    auto out_vec = [&](const string& name, const vec& v) { overwrite(pre + name); put_vector(pre + name, v); };
    out_vec(KEY_NAME, s);
    const char** in_names = f1 ? F1_INPUTS : KART_INPUTS;
    const char** ct_names = f1 ? F1_CONTROLS : KART_CONTROLS;
    int k = 0;
    for (int i = 0; i < n_in; ++i)
    {
        if (string(in_names[i]) == "time") { out_vec("time", time); continue; }
        vec col(np);
        for (int q = 0; q < np; ++q) col[q] = X[(size_t)q * NVo + k];
        out_vec(in_names[i], col);
        ++k;
    }
    const int full_mesh[2] = { 0, f1 ? 2 : 1 };
    for (int c = 0; c < n_ctrl; ++c)
    {
        vec col(np, ctrl_default[c]);
        if (o.aug_kind == FL_AUG_BRAKE_BIAS && c == 3) col = aug_values;          // constant inside every stretch
        for (int fm = 0; fm < 2; ++fm)
            if (full_mesh[fm] == c) for (int q = 0; q < np; ++q) col[q] = X[(size_t)q * NVo + k + fm];
        out_vec(ct_names[c], col);
    }
    out_vec("x", xs); out_vec("y", ys); out_vec("psi", psi);
    for (auto& kv : channels) out_vec(kv.first, kv.second);
    overwrite(pre + "laptime"); scalars[pre + "laptime"] = laptime;
    if (o.aug_kind == FL_AUG_INTEGRAL)                                           // fastestlapc.cpp:1619-1641
    {
        overwrite(pre + "integral_quantities." + o.integral_name);
        scalars[pre + "integral_quantities." + o.integral_name] = aug_values[0];
    }
    overwrite(pre + "optimization_data.number_of_iterations"); scalars[pre + "optimization_data.number_of_iterations"] = (double)iters[0];
    // derivatives with respect to the declared parameters: "<prefix>derivatives/<variable>/<alias>" (fastestlapc.cpp:1585-1615, 1687-1706; the
    // reference fills chassis.velocity.x and time and leaves other requested vectors at zero -- here every input and optimised control
    // gets its derivative); the lap time's by the trapezoid rule over d/dp of dt/ds (optimal_laptime.hpp:664-724)
    for (size_t q = 0; q < dx_dp.size(); ++q)
    {
        const string& alias = sens[q].alias;
        vec dx = dx_dp[q];
        if (!o.closed) dx.insert(dx.begin(), (size_t)NVo, 0.0);          // the given first node does not move (optimal_laptime.hpp:651-659)
        vec d2(np), dtime(np, 0.0);
        for (int i = 0; i < np; ++i)
        {
            const double* x = X.data() + (size_t)i * NVo;
            const double* dxi = dx.data() + (size_t)i * NVo;
            const double mu = nodes[(size_t)i * 6 + 1], ph = nodes[(size_t)i * 6 + 2];
            const double kz = -std::sin(ph) * nodes[(size_t)i * 6 + 4] + std::cos(mu) * std::cos(ph) * nodes[(size_t)i * 6 + 3];
            const double ca = std::cos(x[a_col]), sa = std::sin(x[a_col]), un = x[iu] * ca - x[iv] * sa, r = 1.0 - kz * x[n_col];
            d2[i] = -kz * dxi[n_col] / un - ca * r / (un * un) * dxi[iu] + sa * r / (un * un) * dxi[iv] + r * (x[iv] * ca + x[iu] * sa) / (un * un) * dxi[a_col];
        }
        double dlap = 0.0;
        for (int i = 0; i < (o.closed ? np : np - 1); ++i)
        {
            const int j = (i + 1) % np;
            const double h = (i + 1 < np) ? s[i + 1] - s[i] : trk.length - s[np - 1];
            dlap += h * ((1.0 - o.sigma) * d2[i] + o.sigma * d2[j]);
            if (i + 1 < np) dtime[i + 1] = dlap;
        }
        overwrite(pre + "derivatives/laptime/" + alias); scalars[pre + "derivatives/laptime/" + alias] = dlap;
        int kk = 0;
        for (int i = 0; i < n_in; ++i)
        {
            if (string(in_names[i]) == "time") { out_vec("derivatives/time/" + alias, dtime); continue; }
            vec col(np);
            for (int qn = 0; qn < np; ++qn) col[qn] = dx[(size_t)qn * NVo + kk];
            out_vec(string("derivatives/") + in_names[i] + "/" + alias, col);
            ++kk;
        }
        for (int c = 0; c < n_ctrl; ++c)
        {
            vec col(np, 0.0);
            for (int fm = 0; fm < 2; ++fm)
                if (full_mesh[fm] == c) for (int qn = 0; qn < np; ++qn) col[qn] = dx[(size_t)qn * NVo + kk + fm];
            out_vec(string("derivatives/") + ct_names[c] + "/" + alias, col);
        }
    }
}

// Steady_state::gg_diagram (steady_state.hpp:696-880) at one speed, in the reference's order of solves; accelerations in m/s2
static void do_gg_diagram(const Vehicle& veh, double v, int L, double* ay_out, double* ax_max, double* ax_min)
{
    if (L < 2) throw error("gg_diagram -> at least two points");
    const bool f1 = veh.model == FL_MODEL_F1_3DOF;
    const int ia = f1 ? 11 : 6, iy = f1 ? 12 : 7;
    const double unit = f1 ? G0 : 1.0;                  // acceleration_units
    const std::vector<vec> P = { veh.params };
    // (1) 0 g, (2) maximum lateral acceleration, (3) maximum / minimum longitudinal acceleration at ay = 0
    SteadyResult r0 = steady_batch(veh.model, P, { vec{ v, 0, 0, 0, 0, 0, 0 } }, { steady_guess(veh.model) });
    if (r0.status[0] < 1) throw error("[ERROR] gg_diagram -> steady-state computation did not converge");
    vec x0g = r0.x[0];
    x0g[ia] = 0.0; x0g[iy] = 0.0;
    const SteadyResult rl = steady_batch(veh.model, P, { vec{ v, 0, 0, 1, 1, 0, -1 } }, { x0g });
    if (rl.status[0] < 1) throw error("[ERROR] gg_diagram -> maximum lateral acceleration did not converge");
    const double ay_max = rl.x[0][iy], ax_lat = rl.x[0][ia];
    // bounds: the model's (accelerating / braking variants for the F1, limebeer2014f1.h:62-80), constraint bounds
    vec xl_a, xu_a, xl_b, xu_b, gl, gu;
    if (f1)
    {
        xl_a = { -1.15, -1.15, -1.15, -1.15, 0.0, -2.0, -2.0, -2.0, -2.0, 0.0, -1.0, -8.0, -8.0 }; xu_a = { 0.25, 0.25, 1.00, 1.00, 1.5, 0.1, 0.1, 0.1, 0.1, 1.5, 1.0, 8.0, 8.0 };
        xl_b = { -1.15, -1.15, -1.30, -1.30, 0.0, -2.0, -2.0, -2.0, -2.0, 0.0, -1.0, -8.0, -8.0 }; xu_b = { 0.25, 0.25, 0.25, 0.25, 1.0, 0.1, 0.1, 0.1, 0.1, 1.0, 0.5, 8.0, 8.0 };
        gl = vec(15, 0.0); gu = vec(15, 0.0);
        for (int r = 11; r < 15; ++r) { gl[r] = -1.0; gu[r] = 1.2; }
    }
    else
    {
        xl_a = { -0.5, 1.0e-4, -10 * DEG, -10 * DEG, -10 * DEG, -10 * DEG, -20.0, -20.0 }; xu_a = { 0.5, 0.04, 10 * DEG, 10 * DEG, 10 * DEG, 10 * DEG, 20.0, 20.0 };
        xl_b = xl_a; xu_b = xu_a;
        gl = vec(12, 0.0); gu = vec(12, 0.0);
        for (int r = 6; r < 12; ++r) { gu[r] = r < 8 ? 0.11 : 0.09; gl[r] = -gu[r]; }
    }
    const SteadyResult lmax = steady_batch(veh.model, P, { vec{ v, 0, 0, 1, 0, -1, 0 } }, { x0g }, f1 ? xl_a : vec(), f1 ? xu_a : vec(), f1 ? gl : vec(), f1 ? gu : vec());
    const SteadyResult lmin = steady_batch(veh.model, P, { vec{ v, 0, 0, 1, 0, 1, 0 } }, { x0g }, f1 ? xl_b : vec(), f1 ? xu_b : vec(), f1 ? gl : vec(), f1 ? gu : vec());
    const double ax_lon_max = lmax.x[0][ia], ax_lon_min = lmin.x[0][ia];
    // (4) feasible points of the fan, from the 0 g state; a failure keeps the previous point
    const int nf = L - 1, nvar = (int)x0g.size();
    std::vector<vec> spec(nf), start(nf, x0g);
    vec ay(nf), axf(nf);
    for (int i = 0; i < nf; ++i)
    {
        const double t = (double)i / (double)(L - 1);
        ay[i] = t * ay_max; axf[i] = t * ax_lat;
        spec[i] = { v, axf[i], ay[i], 0, 0, 0, 0 };
    }
    const SteadyResult feas = steady_batch(veh.model, P, spec, start);
    std::vector<vec> xf = feas.x;
    for (int i = 0; i < nf; ++i)
        if (feas.status[i] < 1) { xf[i] = i > 0 ? xf[i - 1] : x0g; axf[i] = i > 0 ? axf[i - 1] : 0.0; }
    // (5) maximum and minimum from that point with the reference's bounds on ax; failures retried from the previous point's solution
    for (int side = 0; side < 2; ++side)
    {
        const double sign = side == 0 ? -1.0 : 1.0;
        vec xl = side == 0 ? xl_a : xl_b, xu = side == 0 ? xu_a : xu_b;
        xl[ia] = side == 0 ? ax_lat - 0.5 / unit : ax_lon_min - 0.1 / unit;
        xu[ia] = side == 0 ? ax_lon_max + 0.1 / unit : ax_lat + 0.1 / unit;
        std::vector<vec> sp(nf), st(nf);
        for (int i = 0; i < nf; ++i)
        {
            sp[i] = { v, 0, ay[i], 1, 0, sign, 0 };
            st[i] = xf[i];
            st[i][ia] = std::min(std::max(axf[i], xl[ia] + 1.0e-6), xu[ia] - 1.0e-6);
            st[i][iy] = ay[i];
        }
        SteadyResult r = steady_batch(veh.model, P, sp, st, xl, xu, gl, gu);
        for (int sweep = 0; sweep < nf; ++sweep)
        {
            std::vector<int> todo;
            for (int i = 1; i < nf; ++i) if (r.status[i] < 1 && r.status[i - 1] >= 1) todo.push_back(i);
            if (todo.empty()) break;
            std::vector<vec> sp2, st2;
            for (int i : todo) { sp2.push_back(sp[i]); vec x = r.x[i - 1]; x[iy] = ay[i]; x[ia] = std::min(std::max(x[ia], xl[ia] + 1.0e-6), xu[ia] - 1.0e-6); st2.push_back(x); }
            const SteadyResult r2 = steady_batch(veh.model, P, sp2, st2, xl, xu, gl, gu);
            bool any = false;
            for (size_t q = 0; q < todo.size(); ++q) if (r2.status[q] >= 1) { r.x[todo[q]] = r2.x[q]; r.status[todo[q]] = r2.status[q]; any = true; }
            if (!any) break;
        }
        double* out = side == 0 ? ax_max : ax_min;
        for (int i = 0; i < nf; ++i)
        {
            if (r.status[i] < 1 && print_level > 0) std::printf("gg_diagram -> point %d of the %s side was not solved\n", i, side == 0 ? "accelerating" : "braking");
            out[i] = r.x[i][ia] * unit;
        }
        out[L - 1] = ax_lat * unit;                       // (6) the last point is the maximum-lateral-acceleration solution
    }
    for (int i = 0; i < nf; ++i) ay_out[i] = ay[i] * unit;
    ay_out[L - 1] = ay_max * unit;
    (void)nvar;
}

} // namespace flc

// ---------------------------------------------------------------------------------------------------------------------
#define FLC_TRY flc::last_error.clear(); try {
#define FLC_CATCH(...) } catch (const std::exception& ex) { flc::last_error = ex.what(); std::printf("[Fastest lap exception] -> %s\n", ex.what()); std::fflush(stdout); return __VA_ARGS__; }

extern "C" {

const char* fastestlapc_last_error(void) { return flc::last_error.c_str(); }

void set_print_level(int level) { flc::print_level = level; }

static std::string describe(const std::string& name)
{
    using namespace flc;
    std::ostringstream os;
    os.precision(17);
    if (vehicles.count(name)) os << vehicles[name].to_xml();
    else if (tracks.count(name)) os << "track \"" << name << "\": " << tracks[name].s.size() << " points, length " << tracks[name].length << (tracks[name].closed ? ", closed" : ", open") << "\n";
    else if (scalars.count(name)) os << scalars[name] << "\n";
    else if (vectors.count(name)) { for (size_t i = 0; i < vectors[name].size(); ++i) os << (i ? ", " : "") << vectors[name][i]; os << "\n"; }
    else throw error("Variable \"" + name + "\" does not exist");
    return os.str();
}

void print_variables(void)
{
    using namespace flc;
    std::printf("Type f1-3dof / kart-6dof:\n");
    for (auto& kv : vehicles) std::printf("    -> %s (%s)\n", kv.first.c_str(), kv.second.model == FL_MODEL_F1_3DOF ? "f1-3dof" : "kart-6dof");
    std::printf("Type track:\n");
    for (auto& kv : tracks) std::printf("    -> %s\n", kv.first.c_str());
    std::printf("Type scalar:\n");
    for (auto& kv : scalars) std::printf("    -> %s\n", kv.first.c_str());
    std::printf("Type vector:\n");
    for (auto& kv : vectors) std::printf("    -> %s\n", kv.first.c_str());
    std::fflush(stdout);
}

void print_variable(const char* name) { FLC_TRY std::printf("%s", describe(name).c_str()); std::fflush(stdout); FLC_CATCH() }

void print_variable_to_string(char* str_out, const int n_char, const char* name)
{
    FLC_TRY
    const std::string text = describe(name);
    if ((int)text.size() + 1 > n_char) throw flc::error("print_variable_to_string -> buffer too small");
    std::memcpy(str_out, text.c_str(), text.size() + 1);
    FLC_CATCH()
}

void create_vehicle_from_xml(const char* name, const char* file)
{
    FLC_TRY
    flc::check_new(name);
    flc::vehicles[name] = flc::vehicle_from_xml(flc::read_file(file));
    FLC_CATCH()
}

void create_vehicle_empty(const char* name, const char* type)
{
    FLC_TRY
    flc::check_new(name);
    flc::Vehicle v;
    if (std::string(type) == "f1-3dof") v.model = FL_MODEL_F1_3DOF;
    else if (std::string(type) == "kart-6dof") v.model = FL_MODEL_KART_6DOF;
    else throw flc::error(std::string("vehicle type \"") + type + "\" not recognized");
    v.params.assign(fl_n_vehicle_params(v.model), 0.0);
    flc::vehicles[name] = v;
    FLC_CATCH()
}

void create_track_from_xml(const char* name, const char* file)
{
    FLC_TRY
    flc::check_new(name);
    flc::tracks[name] = flc::track_from_xml(flc::read_file(file));
    FLC_CATCH()
}

void create_vector(const char* name, const int n, double* data) { FLC_TRY flc::check_new(name); flc::vectors[name] = flc::vec(data, data + n); FLC_CATCH() }
void create_scalar(const char* name, double value) { FLC_TRY flc::check_new(name); flc::scalars[name] = value; FLC_CATCH() }

void copy_variable(const char* old_name, const char* new_name)
{
    FLC_TRY
    using namespace flc;
    check_new(new_name);
    if (vehicles.count(old_name)) vehicles[new_name] = vehicles[old_name];
    else if (tracks.count(old_name)) tracks[new_name] = tracks[old_name];
    else if (scalars.count(old_name)) scalars[new_name] = scalars[old_name];
    else if (vectors.count(old_name)) vectors[new_name] = vectors[old_name];
    else throw error(std::string("Variable \"") + old_name + "\" does not exist");
    FLC_CATCH()
}

void delete_variable(const char* c_name)
{
    FLC_TRY
    using namespace flc;
    const std::string name(c_name);
    if (!name.empty() && name.back() == '*')
    {
        // "prefix*": every variable whose name starts with the prefix (fastestlapc.cpp:300-340)
        const std::string pre = name.substr(0, name.size() - 1);
        auto sweep = [&](auto& table) { for (auto it = table.begin(); it != table.end();) if (it->first.compare(0, pre.size(), pre) == 0) it = table.erase(it); else ++it; };
        sweep(vehicles); sweep(tracks); sweep(scalars); sweep(vectors);
        return;
    }
    if (!(vehicles.erase(name) + tracks.erase(name) + scalars.erase(name) + vectors.erase(name))) throw error("Variable \"" + name + "\" does not exist");
    FLC_CATCH()
}

void move_variable(const char* old_name, const char* new_name)
{
    copy_variable(old_name, new_name);
    if (flc::last_error.empty()) delete_variable(old_name);
}

void variable_type(char* out, const int len, const char* c_name)
{
    FLC_TRY
    using namespace flc;
    const std::string name(c_name);
    std::string t;
    if (vehicles.count(name)) t = vehicles[name].model == FL_MODEL_F1_3DOF ? "f1-3dof" : "kart-6dof";
    else if (tracks.count(name)) t = "track";
    else if (scalars.count(name)) t = "scalar";
    else if (vectors.count(name)) t = "vector";
    else throw error("Variable \"" + name + "\" does not exist");
    if ((int)t.size() + 1 > len) throw error("variable_type -> buffer too small");
    std::memcpy(out, t.c_str(), t.size() + 1);
    FLC_CATCH()
}

double download_scalar(const char* name)
{
    FLC_TRY
    auto it = flc::scalars.find(name);
    if (it == flc::scalars.end()) throw flc::error(std::string("Variable \"") + name + "\" does not exist in the scalar table");
    return it->second;
    FLC_CATCH(0.0)
}

int download_vector_size(const char* name)
{
    FLC_TRY
    auto it = flc::vectors.find(name);
    if (it == flc::vectors.end()) throw flc::error(std::string("Variable \"") + name + "\" does not exist in the vector table");
    return (int)it->second.size();
    FLC_CATCH(0)
}

void download_vector(double* data, const int n, const char* name)
{
    FLC_TRY
    auto it = flc::vectors.find(name);
    if (it == flc::vectors.end()) throw flc::error(std::string("Variable \"") + name + "\" does not exist in the vector table");
    if ((int)it->second.size() != n) throw flc::error("Incorrect input size for variable \"" + std::string(name) + "\". Input: " + std::to_string(n) + ", should be " + std::to_string(it->second.size()));
    std::copy(it->second.begin(), it->second.end(), data);
    FLC_CATCH()
}

void vehicle_type_get_sizes(int* n_inputs, int* n_control, int* n_outputs, const char* type)
{
    FLC_TRY
    const std::string t(type);
    if (t == "f1-3dof") { *n_inputs = 14; *n_control = 4; }
    else if (t == "kart-6dof") { *n_inputs = 13; *n_control = 2; }
    else throw flc::error("[ERROR] vehicle type \"" + t + "\" not recognized");
    *n_outputs = 4;
    FLC_CATCH()
}

void vehicle_type_get_names(char* key_name, char* input_names[], char* control_names[], char* output_names[], const int n_char, const char* type)
{
    FLC_TRY
    const std::string t(type);
    const bool f1 = t == "f1-3dof";
    if (!f1 && t != "kart-6dof") throw flc::error("[ERROR] vehicle type \"" + t + "\" not recognized");
    auto put = [&](char* dst, const char* src) { if ((int)std::strlen(src) + 1 > n_char) throw flc::error("vehicle_type_get_names -> buffer too small"); std::strcpy(dst, src); };
    put(key_name, flc::KEY_NAME);
    for (int i = 0; i < (f1 ? 14 : 13); ++i) put(input_names[i], f1 ? flc::F1_INPUTS[i] : flc::KART_INPUTS[i]);
    for (int i = 0; i < (f1 ? 4 : 2); ++i) put(control_names[i], f1 ? flc::F1_CONTROLS[i] : flc::KART_CONTROLS[i]);
    for (int i = 0; i < 4; ++i) put(output_names[i], flc::OUTPUTS[i]);
    FLC_CATCH()
}

double vehicle_get_output(const char* vname, const double* inputs, const double* controls, const double s, const char* property)
{
    // one output channel of the vehicle at (inputs, controls), at arclength s of its track (fastestlapc.cpp vehicle_get_output):
    // evaluated on the GPU as the fixed first node of a two-node open section
    FLC_TRY
    flc::Vehicle veh = flc::vehicle(vname);
    if (veh.track.empty()) throw flc::error("[ERROR] vehicle_get_output -> the vehicle has no track: call vehicle_change_track first");
    flc::Track& trk = flc::track(veh.track);
    const bool f1 = veh.model == FL_MODEL_F1_3DOF;
    if (!veh.variable.empty()) veh.params = veh.params_at(s);
    if (f1) veh.params[18] = controls[3];          // the brake bias is a parameter of the node program when it is not optimised
    const double ds = std::min(1.0, 0.5 * (trk.length - s));
    const flc::vec mesh = { s, s + ds };
    flc::vec nodes(12), wl(2), wr(2);
    const char* keys[6] = { "yaw", "pitch", "roll", "yaw_dot", "dpitch", "droll" };
    for (int i = 0; i < 2; ++i)
    {
        for (int c = 0; c < 6; ++c) nodes[(size_t)i * 6 + c] = trk.at(keys[c], mesh[i]);
        wl[i] = trk.at("nl", mesh[i]); wr[i] = trk.at("nr", mesh[i]);
    }
    const double du0[2] = { 0.0, 0.0 };
    fl_problem_desc d;
    std::memset(&d, 0, sizeof d);
    d.model = veh.model; d.form = f1 ? FL_FORM_DIRECT : FL_FORM_DERIVATIVE; d.closed = 0; d.elevation = trk.has_elevation() ? 1 : 0; d.n_points = 2;
    d.s = mesh.data(); d.track_length = trk.length; d.track_nodes = nodes.data(); d.wl = wl.data(); d.wr = wr.data();
    d.batch = 1; d.vehicle_params = veh.params.data(); d.sigma = 0.5; d.q0 = inputs; d.u0 = controls; d.du0 = du0; d.device = 0;
    d.wind[0] = veh.wind[0]; d.wind[1] = veh.wind[1];
    fl_nlp* nlp = fl_nlp_create(&d);
    if (!nlp) throw flc::error("[ERROR] vehicle_get_output -> " + flc::fl_error());
    int n = 0, m = 0, which = -1;
    fl_nlp_sizes(nlp, &n, &m, nullptr, nullptr);
    const int n_out = fl_nlp_n_outputs(nlp);
    for (int k = 0; k < n_out; ++k) if (std::string(property) == fl_nlp_output_name(nlp, k)) which = k;
    double value = 0.0;
    std::string msg;
    if (which < 0) msg = std::string("output \"") + property + "\" is not offered";
    else
    {
        // variables of the free node: the same inputs and controls
        const int n_in = f1 ? 14 : 13;
        flc::vec x;
        for (int j = 0; j < n_in; ++j) if (j != n_in - 3) x.push_back(inputs[j]);
        x.push_back(controls[0]); x.push_back(controls[f1 ? 2 : 1]);
        x.resize(n, 0.0);
        flc::vec out((size_t)n_out * 2);
        if (!fl_eval_outputs(nlp, x.data(), out.data())) msg = flc::fl_error();
        else value = out[(size_t)which * 2];
    }
    fl_nlp_destroy(nlp);
    if (!msg.empty()) throw flc::error("[ERROR] vehicle_get_output -> " + msg);
    return value;
    FLC_CATCH(0.0)
}

void vehicle_save_as_xml(const char* name, const char* file)
{
    FLC_TRY
    std::ofstream f(file);
    if (!f) throw flc::error(std::string("cannot write \"") + file + "\"");
    f << flc::vehicle(name).to_xml();
    FLC_CATCH()
}

int track_download_number_of_points(const char* name) { FLC_TRY return (int)flc::track(name).s.size(); FLC_CATCH(0) }
double track_download_length(const char* name) { FLC_TRY return flc::track(name).length; FLC_CATCH(0.0) }

void track_download_data(double* data, const char* name, const int n, const char* variable)
{
    FLC_TRY
    const flc::Track& t = flc::track(name);
    if ((int)t.s.size() != n) throw flc::error("[ERROR] track_download_data -> incorrect number of points");
    const std::string v(variable);
    const std::map<std::string, std::string> keys = { { "centerline.x", "x" }, { "centerline.y", "y" }, { "heading-angle", "yaw" }, { "curvature", "yaw_dot" },
                                                      { "distance-left-boundary", "nl" }, { "distance-right-boundary", "nr" } };
    auto side = [&](const char* w, double sx, double sy) {
        // track boundaries (fastestlapc.cpp track_download_data: left.x ... right.y): centreline -+ n (sin(theta), -cos(theta))
        for (int i = 0; i < n; ++i)
        {
            const double th = t.a.at("yaw")[i], d = t.a.at(w)[i];
            data[i] = (sx != 0.0 ? t.a.at("x")[i] + sx * d * std::sin(th) : t.a.at("y")[i] + sy * d * std::cos(th));
        }
    };
    if (v == "arclength") std::copy(t.s.begin(), t.s.end(), data);
    else if (keys.count(v)) std::copy(t.a.at(keys.at(v)).begin(), t.a.at(keys.at(v)).end(), data);
    else if (v == "left.x") side("nl", 1.0, 0.0);
    else if (v == "left.y") side("nl", 0.0, -1.0);
    else if (v == "right.x") side("nr", -1.0, 0.0);
    else if (v == "right.y") side("nr", 0.0, 1.0);
    else throw flc::error("[ERROR] track_download_data -> variable_name \"" + v + "\" is not valid");
    FLC_CATCH()
}

void vehicle_set_parameter(const char* name, const char* parameter, const double value) { FLC_TRY flc::vehicle(name).set(parameter, value); FLC_CATCH() }

void vehicle_declare_new_constant_parameter(const char* name, const char* path, const char* alias, const double value)
{
    // the value is set, and the alias names the derivatives <compute_sensitivity> of optimal_laptime stores (fastestlapc.cpp:899-939)
    FLC_TRY
        flc::Vehicle& v = flc::vehicle(name);
        v.set(path, value);
        const int slot = v.slot(path);
        if (slot < 0) throw flc::error(std::string("Parameter \"") + path + "\" cannot be differentiated");
        const std::string a = flc::trim(alias ? alias : "");
        for (auto& d : v.declared) if (d.first == a) throw flc::error("Parameter alias \"" + a + "\" already declared");
        v.declared.emplace_back(a, slot);
    FLC_CATCH()
}

void vehicle_declare_new_variable_parameter(const char* name, const char* path, const char* aliases, const int n_parameters, const double* values,
                                            const int mesh_size, const int* mesh_indexes, const double* mesh_points)
{
    FLC_TRY
    flc::Vehicle& v = flc::vehicle(name);
    const int k = v.slot(path);
    if (k < 0) throw flc::error("the wind cannot vary along the lap");
    if (n_parameters < 1) throw flc::error("vehicle_declare_new_variable_parameter -> no values");
    flc::VarParam p;
    p.values.assign(values, values + n_parameters);
    p.idx.assign(mesh_indexes, mesh_indexes + mesh_size);
    p.pts.assign(mesh_points, mesh_points + mesh_size);
    for (int i : p.idx) if (i < 0 || i >= n_parameters) throw flc::error("mesh indexes must address the parameter values");
    if (!std::is_sorted(p.pts.begin(), p.pts.end())) throw flc::error("mesh points must be sorted");
    // "a; b; c": one alias per value (fastestlapc.cpp:962-969), the names of the derivatives of <compute_sensitivity>
    if (aliases && *aliases)
    {
        std::string rest = aliases;
        for (;;)
        {
            const size_t semi = rest.find(';');
            p.aliases.push_back(flc::trim(rest.substr(0, semi)));
            if (semi == std::string::npos) break;
            rest = rest.substr(semi + 1);
        }
        if ((int)p.aliases.size() != n_parameters) throw flc::error("vehicle_declare_new_variable_parameter -> one alias per parameter value");
    }
    v.variable[k] = p;
    v.params[k] = p.values.front();
    FLC_CATCH()
}

void vehicle_change_track(const char* vname, const char* tname) { FLC_TRY flc::track(tname); flc::vehicle(vname).track = tname; FLC_CATCH() }

void track_set_track_limit_correction(const char* name, const char* side, const int n_points, const double* s, const double* w)
{
    // adds a piecewise-linear correction w(s) to the left / right track limit at the track's samples (fastestlapc.cpp:1046-1109)
    FLC_TRY
    flc::Track& t = flc::track(name);
    const std::string sd(side);
    if (sd != "left" && sd != "right") throw flc::error("track_set_track_limit_correction -> side must be \"left\" or \"right\"");
    if (n_points < 1) throw flc::error("track_set_track_limit_correction -> no points");
    flc::vec& lim = t.a[sd == "left" ? "nl" : "nr"];
    for (size_t i = 0; i < t.s.size(); ++i)
    {
        const double sq = t.s[i];
        double c;
        if (sq <= s[0]) c = w[0];
        else if (sq >= s[n_points - 1]) c = w[n_points - 1];
        else { const int k = (int)(std::upper_bound(s, s + n_points, sq) - s); c = w[k - 1] + (sq - s[k - 1]) / (s[k] - s[k - 1]) * (w[k] - w[k - 1]); }
        lim[i] += c;
    }
    FLC_CATCH()
}

void steady_state(double* input_states, double* controls, const char* name, double v, double ax, double ay)
{
    // Steady_state::solve(v, ax, ay) (steady_state.hpp:25-126); ax, ay in m/s2.  Outputs: the model's inputs and controls.
    FLC_TRY
    const flc::Vehicle& veh = flc::vehicle(name);
    const bool f1 = veh.model == FL_MODEL_F1_3DOF;
    const double unit = f1 ? flc::G0 : 1.0;
    const flc::SteadyResult r0 = flc::steady_batch(veh.model, { veh.params }, { flc::vec{ v, 0, 0, 0, 0, 0, 0 } }, { flc::steady_guess(veh.model) });
    flc::SteadyResult r = r0;
    if (ax != 0.0 || ay != 0.0) r = flc::steady_batch(veh.model, { veh.params }, { flc::vec{ v, ax / unit, ay / unit, 0, 0, 0, 0 } }, { r0.x[0] });
    if (r.status[0] < 1) throw flc::error("[ERROR] steady_state -> the computation did not converge");
    const flc::vec& x = r.x[0];
    if (f1)
    {
        const double ang = 10.0 * flc::DEG, psi = x[4] * ang;
        const double q[14] = { x[0], x[1], x[2], x[3], v * std::cos(psi), -v * std::sin(psi), ay / v, x[5], x[6], x[7], x[8], 0.0, 0.0, psi };
        std::copy(q, q + 14, input_states);
        controls[0] = x[9] * ang; controls[1] = 0.0; controls[2] = x[10]; controls[3] = veh.params[18];
    }
    else
    {
        const double psi = x[4];
        const double q[13] = { (x[0] + 1.0) * v / 0.139, v * std::cos(psi), -v * std::sin(psi), ay / v, x[1], x[2], x[3], 0.0, 0.0, 0.0, 0.0, 0.0, psi };
        std::copy(q, q + 13, input_states);
        controls[0] = x[5]; controls[1] = 0.0;
    }
    FLC_CATCH()
}

void propagate_vehicle(double*, double*, const char*, const char*, double, double, double*, int, const char*)
{
    FLC_TRY throw flc::error("propagate_vehicle -> time propagation is not part of the B200 path"); FLC_CATCH()
}

void gg_diagram(double* ay, double* ax_max, double* ax_min, const char* name, double v, const int n_points)
{
    FLC_TRY flc::do_gg_diagram(flc::vehicle(name), v, n_points, ay, ax_max, ax_min); FLC_CATCH()
}

void optimal_laptime(const char* vname, const char* tname, const int n_points, const double* s, const char* options)
{
    FLC_TRY flc::do_optimal_laptime(vname, tname, flc::vec(s, s + n_points), options ? options : ""); FLC_CATCH()
}

void circuit_preprocessor(const char*) { FLC_TRY throw flc::error("circuit_preprocessor -> not part of the B200 path (tracks are read from their XML files)"); FLC_CATCH() }

} // extern "C"
