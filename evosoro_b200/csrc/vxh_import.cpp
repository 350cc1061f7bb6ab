// vxh_import.cpp -- host side above the engine's C-ABI (include/vxhost.h): .vxa import into the flat
// vxg_robot_desc and the fp64 replay of the CLI clock.  Plain host C++ (no CUDA calls), linked into
// libvxgpu.so.  Nothing here simulates; it only builds the model the kernels consume.
//
// What it does follows the reference's loader/importer (file:line cited per function); the code is
// our own: a small DOM parser instead of tinyxml/XML_Rip, and class de-duplication instead of
// per-voxel/per-bond objects.
#include <zlib.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/vxhost.h"

namespace {

thread_local std::string h_err;
int hfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    h_err = buf;
    return code;
}

// ---------------------------------------------------------------------------------------------------
// Minimal XML DOM: elements, attributes, text and CDATA; comments / declarations / DOCTYPE skipped.
// ---------------------------------------------------------------------------------------------------
// base64 as the reference decodes it (CVXC_Structure::FromBase64, VX_Object.cpp:2357: stops at '=' or at the first character
// outside the alphabet; white space inside the element text therefore ends the data, as it does there)
std::string from_base64(const std::string& in) {
    static const char* abc = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    int val[256]; for (int i = 0; i < 256; i++) val[i] = -1;
    for (int i = 0; i < 64; i++) val[(unsigned char)abc[i]] = i;
    std::string out;
    unsigned int acc = 0; int bits = 0;
    for (unsigned char c : in) {
        if (c == '=' || val[c] < 0) break;
        acc = (acc << 6) | (unsigned int)val[c]; bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((char)((acc >> bits) & 0xFF)); }
    }
    return out;
}

struct XNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::string text;          // concatenated text + CDATA children
    bool has_text = false;
    std::vector<std::unique_ptr<XNode>> kids;
    const XNode* child(const char* tag) const {
        for (auto& k : kids) if (k->name == tag) return k.get();
        return nullptr;
    }
    std::vector<const XNode*> children(const char* tag) const {
        std::vector<const XNode*> r;
        for (auto& k : kids) if (k->name == tag) r.push_back(k.get());
        return r;
    }
    const char* attr(const char* key) const {
        for (auto& a : attrs) if (a.first == key) return a.second.c_str();
        return "";
    }
};

struct XParser {
    const char* p; const char* end; std::string err;
    explicit XParser(const std::string& s) : p(s.data()), end(s.data() + s.size()) {}
    bool starts(const char* s) const { size_t n = strlen(s); return (size_t)(end - p) >= n && !memcmp(p, s, n); }
    void skip_ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool skip_until(const char* s) {
        size_t n = strlen(s);
        while ((size_t)(end - p) >= n) { if (!memcmp(p, s, n)) { p += n; return true; } p++; }
        return false;
    }
    static void append_decoded(std::string& out, const char* a, const char* b) {
        while (a < b) {
            if (*a == '&') {
                if (b - a >= 4 && !memcmp(a, "&lt;", 4)) { out += '<'; a += 4; continue; }
                if (b - a >= 4 && !memcmp(a, "&gt;", 4)) { out += '>'; a += 4; continue; }
                if (b - a >= 5 && !memcmp(a, "&amp;", 5)) { out += '&'; a += 5; continue; }
                if (b - a >= 6 && !memcmp(a, "&quot;", 6)) { out += '"'; a += 6; continue; }
                if (b - a >= 6 && !memcmp(a, "&apos;", 6)) { out += '\''; a += 6; continue; }
            }
            out += *a++;
        }
    }
    bool skip_misc() {   // whitespace, comments, processing instructions, DOCTYPE
        for (;;) {
            skip_ws();
            if (starts("<?")) { if (!skip_until("?>")) return false; }
            else if (starts("<!--")) { if (!skip_until("-->")) return false; }
            else if (starts("<!DOCTYPE")) { if (!skip_until(">")) return false; }
            else return true;
        }
    }
    std::unique_ptr<XNode> element() {
        if (p >= end || *p != '<') { err = "expected '<'"; return nullptr; }
        p++;
        const char* s = p;
        while (p < end && !strchr(" \t\r\n/>", *p)) p++;
        std::unique_ptr<XNode> n(new XNode);
        n->name.assign(s, p);
        if (n->name.empty()) { err = "empty tag name"; return nullptr; }
        for (;;) {   // attributes
            skip_ws();
            if (p >= end) { err = "unterminated tag"; return nullptr; }
            if (*p == '/') { if (p + 1 < end && p[1] == '>') { p += 2; return n; } err = "bad '/'"; return nullptr; }
            if (*p == '>') { p++; break; }
            const char* k = p;
            while (p < end && *p != '=' && !strchr(" \t\r\n/>", *p)) p++;
            std::string key(k, p);
            skip_ws();
            if (p >= end || *p != '=') { err = "attribute without value"; return nullptr; }
            p++; skip_ws();
            if (p >= end || (*p != '"' && *p != '\'')) { err = "unquoted attribute"; return nullptr; }
            char q = *p++;
            const char* v = p;
            while (p < end && *p != q) p++;
            if (p >= end) { err = "unterminated attribute"; return nullptr; }
            std::string val; append_decoded(val, v, p);
            p++;
            n->attrs.emplace_back(key, val);
        }
        for (;;) {   // content
            if (p >= end) { err = "unterminated element <" + n->name + ">"; return nullptr; }
            if (starts("<![CDATA[")) {
                p += 9; const char* s2 = p;
                if (!skip_until("]]>")) { err = "unterminated CDATA"; return nullptr; }
                n->text.append(s2, p - 3); n->has_text = true;
            } else if (starts("<!--")) { if (!skip_until("-->")) { err = "unterminated comment"; return nullptr; } }
            else if (starts("<?")) { if (!skip_until("?>")) { err = "unterminated PI"; return nullptr; } }
            else if (starts("</")) {
                p += 2; const char* s2 = p;
                while (p < end && *p != '>') p++;
                std::string close(s2, p);
                while (!close.empty() && strchr(" \t\r\n", close.back())) close.pop_back();
                if (p >= end || close != n->name) { err = "mismatched </" + close + "> for <" + n->name + ">"; return nullptr; }
                p++;
                return n;
            } else if (*p == '<') {
                auto k = element();
                if (!k) return nullptr;
                n->kids.push_back(std::move(k));
            } else {
                const char* s2 = p;
                while (p < end && *p != '<') p++;
                const char* a = s2; const char* b = p;   // trim like a condensing parser
                while (a < b && strchr(" \t\r\n", *a)) a++;
                while (b > a && strchr(" \t\r\n", b[-1])) b--;
                if (b > a) { append_decoded(n->text, a, b); n->has_text = true; }
            }
        }
    }
    std::unique_ptr<XNode> document() {
        if (!skip_misc()) { err = "unterminated prolog"; return nullptr; }
        return element();
    }
};

// FindLoadElement semantics (Utils/XML_Rip.h:73-77): value = atof / atoi of the element's text; an
// element that exists but has no text counts as "not found".
bool load(const XNode* n, const char* tag, double* v) { const XNode* c = n ? n->child(tag) : nullptr; if (!c || !c->has_text) return false; *v = atof(c->text.c_str()); return true; }
bool load(const XNode* n, const char* tag, int* v) { const XNode* c = n ? n->child(tag) : nullptr; if (!c || !c->has_text) return false; *v = atoi(c->text.c_str()); return true; }
bool load(const XNode* n, const char* tag, bool* v) { int i; if (!load(n, tag, &i)) return false; *v = i != 0; return true; }
bool load(const XNode* n, const char* tag, std::string* v) { const XNode* c = n ? n->child(tag) : nullptr; if (!c || !c->has_text) return false; *v = c->text; return true; }

struct Material {   // CVXC_Material, VX_Object.cpp:1344-1441
    int mat_type = 0, mat_model = 0;
    double e = 0, density = 0, poisson = 0, cte = 0, u_static = 0, u_dynamic = 0;
};

}  // namespace

struct vxh_model {
    vxh_info info{};
    int variant = 0;
    // flattened model (owned)
    std::vector<double> nom_pos;
    std::vector<int32_t> vox_class, bond_v1, bond_v2, bond_axis, bond_class, cell_index;
    std::vector<vxg_vox_class> vclasses;
    std::vector<vxg_bond_class> bclasses;
    vxg_robot_desc desc{};
    double temp_period = 0;
};

namespace {

struct Region {     // CVX_FRegion (VX_FRegion.cpp:165-221): one primitive + what it imposes on the voxels it touches
    int prim = -1;                       // PRIM_BOX 0, PRIM_CYLINDER 1, PRIM_SPHERE 2, PRIM_MESH 3 (Utils/... VX_FRegion.h)
    double x = 0, y = 0, z = 0, dx = 0, dy = 0, dz = 0, radius = 0;   // in fractions of the workspace
    int dof = 0;
    double force[3] = {0, 0, 0}, torque[3] = {0, 0, 0}, disp[3] = {0, 0, 0}, adisp[3] = {0, 0, 0};
};

struct Parsed {
    // <Simulator>  (VX_Sim.cpp:263-356, VX_SimGA.cpp:205-220)
    double dt_frac = 0.9, bond_damp = 0.1, col_damp = 1.0, slow_damp = 1.0, col_horizon = 2.0;
    bool self_col = false, max_vel = false, blending = false, volume_fx = false, equilibrium = false;
    int stop_type = 0; double stop_value = 0, afterlife = 0, midlife_freeze = 0, init_cm_time = 0;
    double min_temp_fact = 0.1;
    std::string fitness_file; bool write_fitness = true;
    // <Environment>  (VX_Environment.cpp:123-293)
    std::vector<Region> bcs;
    bool grav_on = false, floor_on = false, temp_on = false, vary_temp = false, sticky = false, falling_prohibited = false;
    bool limited_floor = false, sources = false, fluid = false, fluid_in_voxcad = false, needle = false, norm_by_vol = false, save_traces = false;
    double grav_acc = -9.81, temp_base = 25, temp_amplitude = 0, temp_period = 0.1, growth_amp = 0, drag = 0;
    // <VXC>  (VX_Object.cpp:435-490, :1064-1074, :1733-1930, :2495-2513)
    double lat = 0.001, adj[3] = {1, 1, 1}, line_off[2] = {0, 0}, layer_off[2] = {0, 0};
    std::vector<Material> palette;   // [0] = the empty material
    int nx = 1, ny = 1, nz = 1;
    std::vector<int> cells;          // material index per lattice cell, x fastest
    bool has_phase = false, has_tad = false;
    std::vector<double> phase, tad;  // per occupied-cell rank (VX_Object.cpp:1892-1899)
    // development / evolved stiffness blocks (VX_Object.cpp:1910-2322), same layout
    bool has_final_phase = false, has_final_tad = false, has_init_size = false, has_final_size = false, has_growth_time = false,
         has_start_growth = false, has_stiffness = false;
    std::vector<double> final_phase, final_tad, init_size, final_size, growth_time, start_growth, stiffness;
    double min_growth_time = 0, parent_lifetime = 0;
    bool has_parent_lifetime = false, onset_relative = false, termination_relative = false;
    int steps_in_window = 0;
};

int parse_per_voxel(const XNode* st, const char* tag, const Parsed& P, std::vector<double>& out, bool* present) {
    const XNode* blk = st->child(tag);
    *present = blk != nullptr;
    if (!blk) return 0;
    auto layers = blk->children("Layer");
    if ((int)layers.size() < P.nz) return hfail(VXG_ERR_ARG, "<%s> has %d layers, structure has %d", tag, (int)layers.size(), P.nz);
    out.clear();
    for (int z = 0; z < P.nz; z++) {
        const std::string& raw = layers[z]->text;
        std::vector<std::string> parts;   // split(RawData, ',')
        size_t a = 0;
        for (;;) { size_t c = raw.find(',', a); if (c == std::string::npos) { parts.push_back(raw.substr(a)); break; } parts.push_back(raw.substr(a, c - a)); a = c + 1; }
        for (int k = 0; k < P.nx * P.ny; k++) {
            if (P.cells[(size_t)z * P.nx * P.ny + k] > 0) {
                if (k >= (int)parts.size()) return hfail(VXG_ERR_ARG, "<%s> layer %d is shorter than the structure", tag, z);
                out.push_back(atof(parts[k].c_str()));
            }
        }
    }
    return 0;
}

// CVX_FRegion::ReadXML (VX_FRegion.cpp:165-221) for the `count_tag` regions of one container element
int parse_regions(const XNode* box, const char* count_tag, std::vector<Region>& out) {
    int n = 0;
    if (!load(box, count_tag, &n)) n = 0;
    auto regs = box->children("FRegion");
    if (n > (int)regs.size()) return hfail(VXG_ERR_ARG, "<%s> is %d but only %d <FRegion> elements follow", count_tag, n, (int)regs.size());
    for (int i = 0; i < n; i++) {
        const XNode* r = regs[i];
        Region g;
        if (!load(r, "PrimType", &g.prim)) return hfail(VXG_ERR_ARG, "<FRegion> without <PrimType>");
        if (g.prim == 3) return hfail(VXG_UNSUPPORTED_FEATURE, "mesh-shaped boundary-condition regions are not supported");
        if (g.prim < 0 || g.prim > 3) return hfail(VXG_ERR_ARG, "unknown <PrimType> %d", g.prim);
        load(r, "X", &g.x); load(r, "Y", &g.y); load(r, "Z", &g.z);
        load(r, "dX", &g.dx); load(r, "dY", &g.dy); load(r, "dZ", &g.dz); load(r, "Radius", &g.radius);
        int dof;
        if (load(r, "DofFixed", &dof)) g.dof = (int)(signed char)dof & 0x3F;
        else { bool fixed = false; g.dof = (load(r, "Fixed", &fixed) && fixed) ? 0x3F : 0; }   // pre-DofFixed files
        load(r, "ForceX", &g.force[0]); load(r, "ForceY", &g.force[1]); load(r, "ForceZ", &g.force[2]);
        load(r, "TorqueX", &g.torque[0]); load(r, "TorqueY", &g.torque[1]); load(r, "TorqueZ", &g.torque[2]);
        load(r, "DisplaceX", &g.disp[0]); load(r, "DisplaceY", &g.disp[1]); load(r, "DisplaceZ", &g.disp[2]);
        load(r, "AngDisplaceX", &g.adisp[0]); load(r, "AngDisplaceY", &g.adisp[1]); load(r, "AngDisplaceZ", &g.adisp[2]);
        out.push_back(g);
    }
    return 0;
}

// CPrimitive::IsTouching(P, Dist, Envelope) (VX_FRegion.cpp:420-429 box, :531-534 cylinder = IsIn :495-524,
// :601-605 sphere): does the voxel centred at p (half extents h) touch the region, envelope = the workspace
bool region_touches(const Region& g, const double p[3], const double h[3], const double ws[3]) {
    if (g.prim == 0) {
        const double px = p[0] / ws[0], py = p[1] / ws[1], pz = p[2] / ws[2];
        const double hx = h[0] / ws[0], hy = h[1] / ws[1], hz = h[2] / ws[2];
        return px + hx > g.x && px - hx < g.x + g.dx && py + hy > g.y && py - hy < g.y + g.dy && pz + hz > g.z && pz - hz < g.z + g.dz;
    }
    const double cx = g.x * ws[0], cy = g.y * ws[1], cz = g.z * ws[2];
    const double sx = g.dx * ws[0], sy = g.dy * ws[1], sz = g.dz * ws[2];
    const double wmax = ws[0] > ws[1] ? (ws[0] > ws[2] ? ws[0] : ws[2]) : (ws[1] > ws[2] ? ws[1] : ws[2]);
    const double rad = g.radius * wmax;
    if (g.prim == 1) {
        if (g.dx != 0) return p[0] > cx && p[0] < cx + sx && sqrt((p[1] - cy) * (p[1] - cy) + (p[2] - cz) * (p[2] - cz)) < rad;
        if (g.dy != 0) return p[1] > cy && p[1] < cy + sy && sqrt((p[0] - cx) * (p[0] - cx) + (p[2] - cz) * (p[2] - cz)) < rad;
        if (g.dz != 0) return p[2] > cz && p[2] < cz + sz && sqrt((p[0] - cx) * (p[0] - cx) + (p[1] - cy) * (p[1] - cy)) < rad;
        return false;
    }
    const double dist = sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]);   // Dist->Length()
    return sqrt((p[0] - cx) * (p[0] - cx) + (p[1] - cy) * (p[1] - cy) + (p[2] - cz) * (p[2] - cz)) < rad + dist;
}

int parse_vxa(const XNode* root, int variant, Parsed& P) {
    if (root->name != "VXA") return hfail(VXG_ERR_ARG, "root element is <%s>, expected <VXA>", root->name.c_str());
    const bool lw = variant == VXH_VARIANT_LAND_WATER;
    if (const XNode* S = root->child("Simulator")) {
        if (const XNode* n = S->child("Integration")) { if (!load(n, "DtFrac", &P.dt_frac)) P.dt_frac = 0.9; }
        if (const XNode* n = S->child("Damping")) {
            if (!load(n, "BondDampingZ", &P.bond_damp)) P.bond_damp = 0.1;
            if (!load(n, "ColDampingZ", &P.col_damp)) P.col_damp = 1.0;
            if (!load(n, "SlowDampingZ", &P.slow_damp)) P.slow_damp = 1.0;
        }
        if (const XNode* n = S->child("Collisions")) {
            if (!load(n, "SelfColEnabled", &P.self_col)) P.self_col = false;
            if (!load(n, "CollisionHorizon", &P.col_horizon)) P.col_horizon = 2.0;
        }
        if (const XNode* n = S->child("Features")) {
            load(n, "MaxVelLimitEnabled", &P.max_vel);
            load(n, "BlendingEnabled", &P.blending);
            load(n, "VolumeEffectsEnabled", &P.volume_fx);
        }
        if (const XNode* n = S->child("StopCondition")) {
            load(n, "StopConditionType", &P.stop_type);
            load(n, "StopConditionValue", &P.stop_value);
            if (!lw) { load(n, "AfterlifeTime", &P.afterlife); load(n, "MidLifeFreezeTime", &P.midlife_freeze); }
            load(n, "InitCmTime", &P.init_cm_time);
        }
        if (!lw) P.has_parent_lifetime = load(S, "ParentLifetime", &P.parent_lifetime);   // VX_Sim.cpp:348
        if (const XNode* n = S->child("EquilibriumMode")) load(n, "EquilibriumModeEnabled", &P.equilibrium);
        if (const XNode* n = S->child("GA")) {
            if (!load(n, "FitnessFileName", &P.fitness_file)) P.fitness_file = "";
            if (!load(n, "WriteFitnessFile", &P.write_fitness)) P.write_fitness = true;
        }
        if (lw || !load(S, "MinTempFact", &P.min_temp_fact)) P.min_temp_fact = 0.1;   // LW: MIN_TEMP_FACTOR constant, LW/VX_Sim.h:31
    }
    if (const XNode* E = root->child("Environment")) {
        int rc;
        if (const XNode* n = E->child("Boundary_Conditions")) { if ((rc = parse_regions(n, "NumBCs", P.bcs))) return rc; }
        else {   // legacy layout (VX_Environment.cpp:140-169): fixed regions first, then forced ones
            if (const XNode* f = E->child("Fixed_Regions")) { if ((rc = parse_regions(f, "NumFixed", P.bcs))) return rc; }
            if (const XNode* f = E->child("Forced_Regions")) { if ((rc = parse_regions(f, "NumForced", P.bcs))) return rc; }
        }
        if (const XNode* n = E->child("Gravity")) {
            if (!load(n, "GravEnabled", &P.grav_on)) P.grav_on = false;
            if (!load(n, "GravAcc", &P.grav_acc)) P.grav_acc = -9.81;
            if (!load(n, "FloorEnabled", &P.floor_on)) P.floor_on = false;
        }
        if (const XNode* n = E->child("Thermal")) {
            if (!load(n, "TempEnabled", &P.temp_on)) P.temp_on = false;
            if (!load(n, "TempBase", &P.temp_base)) P.temp_base = 25;
            if (!load(n, "TempAmplitude", &P.temp_amplitude)) {   // legacy TempAmp (VX_Environment.cpp:183-186)
                double amp;
                P.temp_amplitude = load(n, "TempAmp", &amp) ? amp - P.temp_base : 0;
            }
            if (!load(n, "VaryTempEnabled", &P.vary_temp)) P.vary_temp = false;
            if (!load(n, "TempPeriod", &P.temp_period)) P.temp_period = 0.1;
        }
        if (!lw) {
            if (!load(E, "GrowthAmplitude", &P.growth_amp)) P.growth_amp = 0;
            if (!load(E, "MinGrowthTime", &P.min_growth_time)) P.min_growth_time = 0;       // VX_Environment.cpp:210
            load(E, "OnsetRelative", &P.onset_relative); load(E, "TerminationRelative", &P.termination_relative);
            load(E, "NumTimeStepsInWindow", &P.steps_in_window);
            load(E, "StickyFloor", &P.sticky);
            load(E, "FallingProhibited", &P.falling_prohibited);
            load(E, "NeedleInHaystack", &P.needle); load(E, "NormDistByVol", &P.norm_by_vol); load(E, "SaveTraces", &P.save_traces);
            double fr; P.limited_floor = load(E, "FloorRadius", &fr);
            if (const XNode* s = E->child("Sources")) P.sources = s->child("Source") != nullptr;
        } else {
            load(E, "FluidEnvironment", &P.fluid);
            load(E, "AggregateDragCoefficient", &P.drag);
        }
        // _voxcad has no fluid environment: it would read such a file as a terrestrial one.  A caller that hands a swimming
        // experiment's file to the wrong fork gets an error instead of a plausible but wrong fitness.
        if (!lw) { double fl = 0; if (load(E, "FluidEnvironment", &fl) && fl != 0) P.fluid_in_voxcad = true; }
    }
    const XNode* O = root->child("VXC");
    if (!O) O = root->child("DMF");
    if (!O) return hfail(VXG_ERR_ARG, "no <VXC> element");
    std::string version = O->attr("Version");
    if (const XNode* n = O->child("Lattice")) {
        if (!load(n, "Lattice_Dim", &P.lat)) P.lat = 0.001;
        if (!load(n, "X_Dim_Adj", &P.adj[0])) P.adj[0] = 1; if (!load(n, "Y_Dim_Adj", &P.adj[1])) P.adj[1] = 1; if (!load(n, "Z_Dim_Adj", &P.adj[2])) P.adj[2] = 1;
        load(n, "X_Line_Offset", &P.line_off[0]); load(n, "Y_Line_Offset", &P.line_off[1]);
        load(n, "X_Layer_Offset", &P.layer_off[0]); load(n, "Y_Layer_Offset", &P.layer_off[1]);
    }
    P.palette.clear(); P.palette.emplace_back();   // index 0 = empty
    if (const XNode* pal = O->child("Palette")) {
        for (const XNode* m : pal->children("Material")) {
            Material M;
            if (!load(m, "MatType", &M.mat_type)) {
                if (m->child("Structure") || m->child("RandIndex1") || m->child("VXC") || m->child("DMF"))
                    return hfail(VXG_UNSUPPORTED_FEATURE, "compound / dithered / external palette materials are not supported");
                M.mat_type = 0;
            }
            if (M.mat_type != 0) return hfail(VXG_UNSUPPORTED_FEATURE, "palette material type %d (only SINGLE materials are supported)", M.mat_type);
            if (const XNode* me = m->child("Mechanical")) {
                if (!load(me, "MatModel", &M.mat_model)) M.mat_model = 0;
                load(me, "Elastic_Mod", &M.e); load(me, "Density", &M.density); load(me, "Poissons_Ratio", &M.poisson);
                load(me, "CTE", &M.cte); load(me, "uStatic", &M.u_static); load(me, "uDynamic", &M.u_dynamic);
            }
            P.palette.push_back(M);
        }
    }
    const XNode* st = O->child("Structure");
    if (!st) return hfail(VXG_ERR_ARG, "no <Structure> element");
    if (!(version == "" || version == "0.93" || version == "0.94")) return hfail(VXG_UNSUPPORTED_FEATURE, "VXC version %s is not supported", version.c_str());
    // layer encodings of CVXC_Structure::ReadXML (VX_Object.cpp:1755-1798): ASCII_READABLE (digit = material + '0': what evosoro
    // writes), BASE64 (the default branch: raw material bytes, base64), ZLIB (the same deflated -- the reference needs
    // USE_ZLIB_COMPRESSION defined to read it, which its shipped Makefile does not do), RAW_DATA (deprecated: bytes as they are)
    std::string comp = st->attr("Compression");
    if (comp == "QT_ZLIB") return hfail(VXG_UNSUPPORTED_FEATURE, "structure compression 'QT_ZLIB' is not supported (deprecated Qt encoding; VX_Object.cpp:1756-1769 refuses it without Qt as well)");
    if (!load(st, "X_Voxels", &P.nx)) P.nx = 1; if (!load(st, "Y_Voxels", &P.ny)) P.ny = 1; if (!load(st, "Z_Voxels", &P.nz)) P.nz = 1;
    if (P.nx <= 0 || P.ny <= 0 || P.nz <= 0 || (long long)P.nx * P.ny * P.nz > 100000000LL) return hfail(VXG_ERR_ARG, "bad structure dimensions");
    const XNode* data = st->child("Data");
    if (!data) return hfail(VXG_ERR_ARG, "no <Data> element");
    auto layers = data->children("Layer");
    if ((int)layers.size() < P.nz) return hfail(VXG_ERR_ARG, "structure has %d layers, expected %d", (int)layers.size(), P.nz);
    P.cells.assign((size_t)P.nx * P.ny * P.nz, 0);
    for (int z = 0; z < P.nz; z++) {
        std::string raw = layers[z]->text;
        const int bias = comp == "ASCII_READABLE" ? 48 : 0;
        if (comp == "ZLIB") {
            const std::string packed = from_base64(raw);
            raw.assign((size_t)P.nx * P.ny, '\0');
            uLongf dest = (uLongf)raw.size();
            if (uncompress((Bytef*)&raw[0], &dest, (const Bytef*)packed.data(), (uLong)packed.size()) != Z_OK) return hfail(VXG_ERR_ARG, "layer %d: ZLIB data does not inflate", z);
            raw.resize(dest);
        } else if (comp != "ASCII_READABLE" && comp != "RAW_DATA") raw = from_base64(raw);   // BASE64 and anything unnamed, as the reference's default branch
        if ((int)raw.size() != P.nx * P.ny) return hfail(VXG_ERR_ARG, "Voxel layer data not present or does not match expected size.");
        for (int k = 0; k < P.nx * P.ny; k++) {
            int m = (int)(signed char)(raw[k] - bias);   // VX_Object.cpp:1789
            if (m < 0 || m >= (int)P.palette.size()) return hfail(VXG_ERR_ARG, "cell (layer %d, %d) refers to material %d outside the palette", z, k, m);
            P.cells[(size_t)z * P.nx * P.ny + k] = m;
        }
    }
    int rc;
    if ((rc = parse_per_voxel(st, "PhaseOffset", P, P.phase, &P.has_phase))) return rc;
    if ((rc = parse_per_voxel(st, "Stiffness", P, P.stiffness, &P.has_stiffness))) return rc;
    if (!lw) {   // the land/water fork reads PhaseOffset and Stiffness only (LW/VX_Object.cpp:1879-1970)
        if ((rc = parse_per_voxel(st, "TempAmpDamp", P, P.tad, &P.has_tad))) return rc;
        if ((rc = parse_per_voxel(st, "FinalPhaseOffset", P, P.final_phase, &P.has_final_phase))) return rc;
        if ((rc = parse_per_voxel(st, "FinalTempAmpDamp", P, P.final_tad, &P.has_final_tad))) return rc;
        if ((rc = parse_per_voxel(st, "InitialVoxelSize", P, P.init_size, &P.has_init_size))) return rc;
        if ((rc = parse_per_voxel(st, "FinalVoxelSize", P, P.final_size, &P.has_final_size))) return rc;
        if ((rc = parse_per_voxel(st, "GrowthTime", P, P.growth_time, &P.has_growth_time))) return rc;
        if ((rc = parse_per_voxel(st, "StartGrowthTime", P, P.start_growth, &P.has_start_growth))) return rc;
    }
    // <StiffnessPlasticityRate>, <KP>, <KI>, <ANTIWINDUP> are stored per voxel by the reference (VX_Sim.cpp:977-981) but no
    // step code reads them: accepted and ignored.  Synapse weights belong to the removed neural-net controller.
    static const char* unsupported_blocks[] = {"SynapseWeights", "Weights"};
    for (const char* t : unsupported_blocks)
        if (st->child(t)) return hfail(VXG_UNSUPPORTED_FEATURE, "per-voxel <%s> is not implemented on the device", t);
    return 0;
}

int check_supported(const Parsed& P, int ignore_collisions) {
    if (P.fluid_in_voxcad) return hfail(VXG_UNSUPPORTED_FEATURE, "<FluidEnvironment>1 is a land/water-fork feature: import the file with VXH_VARIANT_LAND_WATER (voxelyze --lw)");
    if (P.max_vel) return hfail(VXG_UNSUPPORTED_FEATURE, "MaxVelLimitEnabled is not supported");
    if (P.blending) return hfail(VXG_UNSUPPORTED_FEATURE, "BlendingEnabled is not supported");
    if (P.volume_fx) return hfail(VXG_UNSUPPORTED_FEATURE, "VolumeEffectsEnabled is not supported");
    if (P.equilibrium) return hfail(VXG_UNSUPPORTED_FEATURE, "EquilibriumModeEnabled is not supported");
    if (P.steps_in_window > 0) return hfail(VXG_UNSUPPORTED_FEATURE, "NumTimeStepsInWindow > 0 (velocity-adjusted development windows) is not supported");
    if (!P.bcs.empty() && (P.line_off[0] != 0 || P.line_off[1] != 0 || P.layer_off[0] != 0 || P.layer_off[1] != 0))
        return hfail(VXG_UNSUPPORTED_FEATURE, "boundary-condition regions on a lattice with line / layer offsets are not supported");
    if (P.limited_floor) return hfail(VXG_UNSUPPORTED_FEATURE, "FloorRadius (limited floor) is not supported");
    if (P.sources) return hfail(VXG_UNSUPPORTED_FEATURE, "environmental <Sources> are not supported");
    if (P.falling_prohibited) return hfail(VXG_UNSUPPORTED_FEATURE, "FallingProhibited is not supported");
    if (P.needle) return hfail(VXG_UNSUPPORTED_FEATURE, "NeedleInHaystack is not supported");
    if (P.norm_by_vol) return hfail(VXG_UNSUPPORTED_FEATURE, "NormDistByVol (volume traces) is not supported");
    if (P.save_traces) return hfail(VXG_UNSUPPORTED_FEATURE, "SaveTraces (CM traces in the result file) is not supported");
    for (size_t i = 1; i < P.palette.size(); i++)
        if (P.palette[i].mat_model != 0 && P.palette[i].mat_model != 1) {   // MDL_LINEAR, MDL_LINEAR_FAIL
            bool used = false;
            for (int c : P.cells) if (c == (int)i) { used = true; break; }
            if (used) return hfail(VXG_UNSUPPORTED_FEATURE, "material %d uses a non-linear model (MatModel %d)", (int)i, P.palette[i].mat_model);
        }
    return 0;
}

// CVX_Sim::Import (VX_Sim.cpp:488-717) + CVX_Voxel::SetMaterial (VX_Voxel.cpp:94-128) + CVX_Bond::LinkVoxels /
// UpdateConstants (VX_Bond.cpp:65-173) + ResetSimulation's per-voxel parameters (VX_Sim.cpp:878-991) +
// CalcMaxDt (VX_Sim.cpp:1693-1727), flattened into classes.
int flatten(const Parsed& P, int variant, int ignore_collisions, vxh_model& M) {
    const bool lw = variant == VXH_VARIANT_LAND_WATER;
    const int nx = P.nx, ny = P.ny, nz = P.nz;
    const size_t ncell = (size_t)nx * ny * nz;
    std::vector<int> x2s(ncell, -1);
    int n_vox = 0;
    for (size_t i = 0; i < ncell; i++) if (P.cells[i] != 0) x2s[i] = n_vox++;
    if (n_vox == 0) return hfail(VXG_ERR_ARG, "No voxels in object");
    if (P.has_phase && (int)P.phase.size() != n_vox) return hfail(VXG_ERR_ARG, "<PhaseOffset> does not cover every voxel");
    if (P.has_tad && (int)P.tad.size() != n_vox) return hfail(VXG_ERR_ARG, "<TempAmpDamp> does not cover every voxel");
    {
        const std::pair<bool, const std::vector<double>*> blocks[] = {{P.has_final_phase, &P.final_phase}, {P.has_final_tad, &P.final_tad},
            {P.has_init_size, &P.init_size}, {P.has_final_size, &P.final_size}, {P.has_growth_time, &P.growth_time},
            {P.has_start_growth, &P.start_growth}, {P.has_stiffness, &P.stiffness}};
        for (auto& b : blocks) if (b.first && (int)b.second->size() != n_vox) return hfail(VXG_ERR_ARG, "a per-voxel block does not cover every voxel");
    }

    // GetLatDimEnv() = Lattice_Dim * DimAdj (VX_Object.h:377); voxels are forced cubic with the X pitch (VX_Sim.cpp:527)
    const double ld[3] = {P.lat * P.adj[0], P.lat * P.adj[1], P.lat * P.adj[2]};
    const double s = ld[0];

    M.nom_pos.resize((size_t)n_vox * 3); M.vox_class.resize(n_vox); M.cell_index.resize(n_vox);
    typedef std::vector<double> VKey;   // material, phaseOffset, TempAmpDamp, boundary-condition record, development values, evolved stiffness
    std::map<VKey, int> vkeys;

    // Boundary conditions (VX_Sim.cpp:517-582): a region's force / torque is shared equally by the voxels it touches
    // (GetNumTouching, VX_Environment.cpp:330-347); fixed DOFs accumulate, prescribed displacements overwrite per axis.
    struct VoxBC { int dof = 0; double f[3] = {0, 0, 0}, t[3] = {0, 0, 0}, d[3] = {0, 0, 0}, a[3] = {0, 0, 0}; };
    std::vector<VoxBC> bc_recs(1);                   // record 0 = unconstrained
    std::map<std::vector<double>, int> bc_ids;
    std::vector<int> vox_bc(n_vox, 0);
    if (!P.bcs.empty()) {
        const double ws[3] = {ld[0] * nx, ld[1] * ny, ld[2] * nz};   // GetWorkSpace (VX_Object.cpp:492-497), no offsets
        const double half[3] = {ld[0] / 2.0, ld[1] / 2.0, ld[2] / 2.0};
        auto cell_pos = [&](size_t i, double p[3]) {
            int iz = (int)(i / ((size_t)nx * ny)), iy = (int)((i - (size_t)iz * nx * ny) / nx), ix = (int)(i - (size_t)iz * nx * ny - (size_t)iy * nx);
            p[0] = ld[0] * (ix + 0.5); p[1] = ld[1] * (iy + 0.5); p[2] = ld[2] * (iz + 0.5);
        };
        // GetNumTouching tests GetXYZ(i) WITH offsets (VX_Object.cpp:513-545): on an offset-free lattice that is the nominal
        // position shifted by the reference's Eps = 1e-6 pitches in x and y, so a region edge that coincides with a voxel face
        // can count a voxel the import loop (nominal position) does not constrain -- restated as is.
        auto count_pos = [&](size_t i, double p[3]) {
            int iz = (int)(i / ((size_t)nx * ny)), iy = (int)((i - (size_t)iz * nx * ny) / nx), ix = (int)(i - (size_t)iz * nx * ny - (size_t)iy * nx);
            const double eps = 0.000001;
            double ox = eps - (int)(eps / P.adj[0]) * P.adj[0], oy = eps - (int)(eps / P.adj[1]) * P.adj[1];
            p[0] = P.lat * (P.adj[0] * (0.5 + ix) + ox); p[1] = P.lat * (P.adj[1] * (0.5 + iy) + oy); p[2] = P.lat * P.adj[2] * (0.5 + iz);
        };
        std::vector<int> sizes(P.bcs.size(), 0);
        for (size_t j = 0; j < P.bcs.size(); j++)
            for (size_t i = 0; i < ncell; i++) { if (x2s[i] < 0) continue; double p[3]; count_pos(i, p); if (region_touches(P.bcs[j], p, half, ws)) sizes[j]++; }
        for (size_t i = 0; i < ncell; i++) {
            if (x2s[i] < 0) continue;
            double p[3]; cell_pos(i, p);
            VoxBC b; bool any = false;
            for (size_t j = 0; j < P.bcs.size(); j++) {
                const Region& g = P.bcs[j];
                if (!region_touches(g, p, half, ws)) continue;
                any = true;
                b.dof |= g.dof;
                for (int a = 0; a < 3; a++) {
                    b.f[a] += g.force[a] / sizes[j]; b.t[a] += g.torque[a] / sizes[j];
                    if (g.dof & (1 << a)) b.d[a] = g.disp[a];
                    if (g.dof & (8 << a)) b.a[a] = g.adisp[a];
                }
            }
            if (!any) continue;
            std::vector<double> key = {(double)b.dof, b.f[0], b.f[1], b.f[2], b.t[0], b.t[1], b.t[2], b.d[0], b.d[1], b.d[2], b.a[0], b.a[1], b.a[2]};
            auto it = bc_ids.find(key);
            if (it == bc_ids.end()) { it = bc_ids.emplace(key, (int)bc_recs.size()).first; bc_recs.push_back(b); }
            vox_bc[x2s[i]] = it->second;
        }
    }
    M.vclasses.clear();
    for (size_t i = 0; i < ncell; i++) {
        int sidx = x2s[i];
        if (sidx < 0) continue;
        int iz = (int)(i / ((size_t)nx * ny)), iy = (int)((i - (size_t)iz * nx * ny) / nx), ix = (int)(i - (size_t)iz * nx * ny - (size_t)iy * nx);
        M.nom_pos[3 * sidx + 0] = ld[0] * (ix + 0.5);   // GetXYZ(.., WithOff=false)  VX_Object.cpp:543
        M.nom_pos[3 * sidx + 1] = ld[1] * (iy + 0.5);
        M.nom_pos[3 * sidx + 2] = ld[2] * (iz + 0.5);
        M.cell_index[sidx] = (int32_t)i;
        int mat = P.cells[i];
        float phase = P.has_phase ? (float)P.phase[sidx] : 0.0f;   // VXS_Voxel.h: `float phaseOffset`
        double tad = P.has_tad ? (double)(float)P.tad[sidx] : 1.0;   // `float TempAmpDamp` (VXS_Voxel.h:94)

        // Per-voxel development parameters as ResetSimulation derives them (VX_Sim.cpp:878-975); the members are floats.
        const double stop_v = P.stop_value, icm = P.init_cm_time, midlife = lw ? 0.0 : P.midlife_freeze;
        const double parent_life = P.has_parent_lifetime ? P.parent_lifetime : stop_v;
        const float onset_bound = (float)(P.onset_relative ? parent_life : stop_v);
        const float term_bound = (float)(P.termination_relative ? parent_life : stop_v);
        float init_size_f = (float)s, final_size_f, start_growth_f, growth_time_f;
        int dev_flags = 0;
        if (P.has_init_size) {
            double f = 1 + (P.growth_amp * P.init_size[sidx]);
            double eff = (f < P.min_temp_fact) ? P.min_temp_fact : f;
            init_size_f = (float)(eff * s);
        }
        if (P.has_final_size) {
            double f = 1 + (P.growth_amp * P.final_size[sidx]);
            double eff = (f < P.min_temp_fact) ? P.min_temp_fact : f;
            final_size_f = (float)(eff * s);
            dev_flags |= VXG_DEV_FINAL_SIZE;
        } else final_size_f = init_size_f;
        if (P.has_start_growth) {
            double g = P.start_growth[sidx] * (onset_bound - icm) + icm;
            if (g >= onset_bound - P.min_growth_time) start_growth_f = (float)(onset_bound - P.min_growth_time);
            else start_growth_f = (float)g;
        } else if (P.has_final_size || P.has_growth_time) start_growth_f = (float)icm;
        else start_growth_f = (float)(stop_v - midlife);
        if (P.has_growth_time) {
            double g = P.growth_time[sidx] * (term_bound - start_growth_f - midlife);
            growth_time_f = (float)((g <= P.min_growth_time) ? P.min_growth_time : g);
        } else if (P.has_final_size) growth_time_f = (float)(term_bound - start_growth_f - midlife);
        else growth_time_f = (float)P.min_growth_time;
        float final_phase_f = P.has_final_phase ? (float)P.final_phase[sidx] : 0.0f;
        float final_tad_f = P.has_final_tad ? (float)P.final_tad[sidx] : 1.0f;
        if (P.has_final_phase) dev_flags |= VXG_DEV_FINAL_PHASE;
        if (P.has_final_tad) dev_flags |= VXG_DEV_FINAL_TAD;
        if (P.has_init_size || P.has_final_size) dev_flags |= VXG_DEV_SIZE_ADJUST;
        if (lw) dev_flags = 0;
        const double stiff = P.has_stiffness ? P.stiffness[sidx] : -1.0;

        VKey key = {(double)mat, (double)phase, tad, (double)vox_bc[sidx], (double)init_size_f, (double)final_size_f, (double)start_growth_f,
                    (double)growth_time_f, (double)final_phase_f, (double)final_tad_f, (double)dev_flags, stiff};
        auto it = vkeys.find(key);
        if (it == vkeys.end()) {
            const Material& pm = P.palette[mat];
            vxg_vox_class c; memset(&c, 0, sizeof c);
            double volume = s * s * s;
            c.mass = volume * pm.density;
            c.inertia = c.mass * (s * s) / 6;
            if (volume == 0 || c.mass == 0 || c.inertia == 0) return hfail(VXG_ERR_ARG, "material %d gives a voxel of zero mass", mat);
            c.mass_inv = 1 / c.mass; c.inertia_inv = 1 / c.inertia;
            c.damp_lin = 2 * sqrt(c.mass * pm.e * s);
            c.damp_ang = 2 * sqrt(c.inertia * pm.e * s * s * s);
            c.k_floor = 2 * pm.e * s;
            c.mu_static = pm.u_static; c.mu_dynamic = pm.u_dynamic;
            c.nom_size = s; c.cte = pm.cte; c.temp_amp_damp = tad; c.phase_offset = (double)phase; c.e_mod = pm.e;
            c.init_size = lw ? s : (double)init_size_f;   // `float initialVoxelSize` (VXS_Voxel.h:100, VX_Sim.cpp:905)
            if (!lw) {
                c.final_size = final_size_f; c.growth_time = growth_time_f; c.start_growth_time = start_growth_f;
                c.final_phase_offset = final_phase_f; c.final_temp_amp_damp = final_tad_f; c.dev_flags = dev_flags;
            }
            if (P.has_stiffness) {
                // _voxcad: SetEMod at reset (VX_Sim.cpp:983-988) changes Vox_E AFTER SetMaterial cached the damping terms and after the
                // internal bonds were linked -- only what reads Vox_E later follows it: floor contact (GetLinearStiffness, VX_Voxel.h:64)
                // and contact bonds (LinkVoxels at creation, VX_Bond.cpp:83).  Land/water fork: SetVoxData runs before the bonds are
                // linked (LW/VX_Sim.cpp:541) and its SetEMod refreshes the damping terms (LW/VX_Voxel.h:61): everything follows it.
                c.e_mod = stiff;
                c.k_floor = 2 * stiff * s;
                if (lw) { c.damp_lin = 2 * sqrt(c.mass * stiff * s); c.damp_ang = 2 * sqrt(c.inertia * stiff * s * s * s); }
            }
            const VoxBC& bc = bc_recs[vox_bc[sidx]];
            c.dof_fixed = bc.dof; c.mat_index = mat;
            for (int a = 0; a < 3; a++) { c.ext_force[a] = bc.f[a]; c.ext_torque[a] = bc.t[a]; c.ext_disp[a] = bc.d[a]; c.ext_tdisp[a] = bc.a[a]; }
            it = vkeys.emplace(key, (int)M.vclasses.size()).first;
            M.vclasses.push_back(c);
        }
        M.vox_class[sidx] = it->second;
    }

    // bonds: for each voxel look +X, +Y, +Z (VX_Sim.cpp:621-641)
    M.bond_v1.clear(); M.bond_v2.clear(); M.bond_axis.clear(); M.bond_class.clear(); M.bclasses.clear();
    typedef std::tuple<int, int, long long, double, double> BKey;   // material 1, material 2, pitch in 1e-9 voxel sizes (may differ per axis), linking moduli
    std::map<BKey, int> bkeys;
    double max_freq2 = 0;
    for (int v1 = 0; v1 < n_vox; v1++) {
        size_t i = (size_t)M.cell_index[v1];
        int iz = (int)(i / ((size_t)nx * ny)), iy = (int)((i - (size_t)iz * nx * ny) / nx), ix = (int)(i - (size_t)iz * nx * ny - (size_t)iy * nx);
        for (int ax = 0; ax < 3; ax++) {
            int jx = ix + (ax == 0), jy = iy + (ax == 1), jz = iz + (ax == 2);
            if (jx >= nx || jy >= ny || jz >= nz) continue;
            size_t j = (size_t)jz * nx * ny + (size_t)jy * nx + jx;
            int v2 = x2s[j];
            if (v2 < 0) continue;
            const vxg_vox_class& c1 = M.vclasses[M.vox_class[v1]]; const vxg_vox_class& c2 = M.vclasses[M.vox_class[v2]];
            if ((c1.dof_fixed & 0x3F) == 0x3F && (c2.dof_fixed & 0x3F) == 0x3F) continue;   // CreatePermBond: no bond between two fully fixed voxels (VX_Sim.cpp:728)
            const Material& m1 = P.palette[c1.mat_index]; const Material& m2 = P.palette[c2.mat_index];
            // Vox_E at LinkVoxels time (VX_Bond.cpp:83): the palette modulus, or the evolved one in the land/water fork
            double E1 = lw ? c1.e_mod : m1.e, E2 = lw ? c2.e_mod : m2.e, u1 = m1.poisson, u2 = m2.poisson;
            if (E1 == 0 || E2 == 0) continue;                           // LinkVoxels fails: no bond (VX_Bond.cpp:87)
            if (u1 < 0 || u1 > 0.5 || u2 < 0 || u2 > 0.5) continue;     // (VX_Bond.cpp:88)
            const double pitch = M.nom_pos[3 * v2 + ax] - M.nom_pos[3 * v1 + ax];
            BKey key(c1.mat_index, c2.mat_index, llround(pitch / s * 1e9), E1, E2);   // bond constants depend on the materials / moduli only (not on per-voxel phase)
            auto it = bkeys.find(key);
            if (it == bkeys.end()) {
                vxg_bond_class b; memset(&b, 0, sizeof b);
                double E = (E1 * E2 / (E1 + E2)) * 2;
                double u = (u1 == 0 && u2 == 0) ? 0 : (u1 * u2 / (u1 + u2)) * 2;
                double L = (c1.nom_size + c2.nom_size) * 0.5;
                double G = E / (2 * (1 + u));
                double A = L * L, Iy = L * L * L * L / 12, J = L * L * (L * L + L * L) / 12;
                b.a1 = E * A / L; b.a2 = G * J / L;
                b.b1 = 12 * E * Iy / (L * L * L); b.b2 = 6 * E * Iy / (L * L); b.b3 = 2 * E * Iy / L;
                b.len = L; b.area = c1.nom_size * c1.nom_size;           // CSArea reset to s^2 by both voxels (VXS_Voxel.cpp:625-631)
                b.e1 = lw ? (double)(float)E1 : E1; b.e2 = lw ? (double)(float)E2 : E2;   // LW: `float voxelEmod` (LW/VX_Object.cpp:1472-1474)
                b.nom_dist = pitch;
                double FM1 = c1.mass * c1.nom_size / 2, FM2 = c2.mass * c2.nom_size / 2;
                b.c_a1m1 = 2.0 * sqrt(b.a1 * c1.mass); b.c_a1m2 = 2.0 * sqrt(b.a1 * c2.mass);
                b.c_a2i1 = 2.0 * sqrt(b.a2 * c1.inertia); b.c_a2i2 = 2.0 * sqrt(b.a2 * c2.inertia);
                b.c_b1m1 = 2.0 * sqrt(b.b1 * c1.mass); b.c_b1m2 = 2.0 * sqrt(b.b1 * c2.mass);
                b.c_b2fm1 = 2.0 * sqrt(b.b2 * FM1); b.c_b2fm2 = 2.0 * sqrt(b.b2 * FM2);
                b.c_b3i1 = 2.0 * sqrt(b.b3 * c1.inertia); b.c_b3i2 = 2.0 * sqrt(b.b3 * c2.inertia);
                b.homogeneous = (c1.mat_index == c2.mat_index) && (!lw || E1 == E2) ? 1 : 0;   // VX_Bond.cpp:75; LW also compares E
                it = bkeys.emplace(key, (int)M.bclasses.size()).first;
                M.bclasses.push_back(b);
            }
            const vxg_bond_class& b = M.bclasses[it->second];
            if (b.a1 / c1.mass > max_freq2) max_freq2 = b.a1 / c1.mass;   // CalcMaxDt (VX_Sim.cpp:1709-1710)
            if (b.a1 / c2.mass > max_freq2) max_freq2 = b.a1 / c2.mass;
            M.bond_v1.push_back(v1); M.bond_v2.push_back(v2); M.bond_axis.push_back(ax); M.bond_class.push_back(it->second);
        }
    }
    if (M.bond_v1.empty())   // no bonds: E / mass of the voxels (VX_Sim.cpp:1714-1721)
        for (int v = 0; v < n_vox; v++) { const vxg_vox_class& c = M.vclasses[M.vox_class[v]]; if (c.e_mod / c.mass > max_freq2) max_freq2 = c.e_mod / c.mass; }
    double max_freq = sqrt(max_freq2);
    double optimal_dt = 1.0 / (max_freq * 2 * 3.1415926);

    vxg_sim_params& sp = M.desc.sim; memset(&sp, 0, sizeof sp);
    sp.dt = P.dt_frac * optimal_dt;
    sp.bond_damp_z = P.bond_damp; sp.col_damp_z = P.col_damp; sp.slow_damp_z = P.slow_damp;
    sp.grav_acc = P.grav_on ? P.grav_acc : 0.0;
    sp.init_cm_time = P.init_cm_time;
    sp.temp_period = (double)(float)P.temp_period;        // `float TempPeriod`, `float TempAmplitude` per voxel (VXS_Voxel.h:92-111)
    sp.temp_amplitude = (double)(float)P.temp_amplitude;
    sp.min_scale_fact = P.min_temp_fact;
    sp.max_scale_fact = 1.0 + P.growth_amp;
    sp.growth_amplitude = lw ? 0.0 : P.growth_amp; sp.midlife_freeze_time = lw ? 0.0 : P.midlife_freeze; sp.stop_condition_value = lw ? 0.0 : P.stop_value;   // read by the _voxcad scale law only
    sp.lattice_dim = P.lat; sp.drag_coeff = P.drag; sp.col_horizon = P.col_horizon;
    uint32_t fl = 0;
    if (P.floor_on) fl |= VXG_F_FLOOR;
    if (P.temp_on) fl |= VXG_F_TEMPERATURE;
    if (P.sticky) fl |= VXG_F_STICKY_FLOOR;
    if (lw) { fl |= VXG_F_LW_SCALE; if (P.fluid) fl |= VXG_F_FLUID; }
    if (P.self_col && !ignore_collisions) fl |= VXG_F_COLLISIONS;
    sp.flags = fl;

    vxg_robot_desc& d = M.desc;
    d.n_vox = n_vox; d.n_bond = (int)M.bond_v1.size(); d.n_vclass = (int)M.vclasses.size(); d.n_bclass = (int)M.bclasses.size();
    d.nom_pos = M.nom_pos.data(); d.vox_class = M.vox_class.data();
    d.bond_v1 = M.bond_v1.data(); d.bond_v2 = M.bond_v2.data(); d.bond_axis = M.bond_axis.data(); d.bond_class = M.bond_class.data();
    d.vclasses = M.vclasses.data(); d.bclasses = M.bclasses.data();

    vxh_info& I = M.info; memset(&I, 0, sizeof I);
    I.n_vox = n_vox; I.n_bond = d.n_bond; I.x_voxels = nx; I.y_voxels = ny; I.z_voxels = nz;
    I.stop_condition_type = P.stop_type; I.self_col_enabled = P.self_col ? 1 : 0; I.write_fitness_file = P.write_fitness ? 1 : 0;
    I.stop_condition_value = P.stop_value; I.init_cm_time = P.init_cm_time; I.afterlife_time = P.afterlife;
    I.dt_frac = P.dt_frac; I.optimal_dt = optimal_dt; I.lattice_dim = P.lat; I.col_horizon = P.col_horizon;
    snprintf(I.fitness_file_name, sizeof I.fitness_file_name, "%s", P.fitness_file.c_str());
    M.variant = variant; M.temp_period = P.temp_period;
    return 0;
}

int load_text(const std::string& text, int variant, int ignore_collisions, vxh_model** out) {
    if (!out) return hfail(VXG_ERR_ARG, "null output pointer");
    *out = nullptr;
    if (variant != VXH_VARIANT_VOXCAD && variant != VXH_VARIANT_LAND_WATER) return hfail(VXG_ERR_ARG, "unknown engine variant %d", variant);
    XParser xp(text);
    std::unique_ptr<XNode> root = xp.document();
    if (!root) return hfail(VXG_ERR_ARG, "Xml read error: %s", xp.err.c_str());
    Parsed P;
    int rc = parse_vxa(root.get(), variant, P);
    if (rc) return rc;
    if ((rc = check_supported(P, ignore_collisions))) return rc;
    std::unique_ptr<vxh_model> M(new vxh_model);
    if ((rc = flatten(P, variant, ignore_collisions, *M))) return rc;
    *out = M.release();
    return 0;
}

}  // namespace

extern "C" const char* vxh_last_error(void) { return h_err.c_str(); }

extern "C" int vxh_load_vxa_string(const char* text, int variant, int ignore_collisions, vxh_model** out) {
    if (!text) return hfail(VXG_ERR_ARG, "null text");
    return load_text(std::string(text), variant, ignore_collisions, out);
}

extern "C" int vxh_load_vxa_file(const char* path, int variant, int ignore_collisions, vxh_model** out) {
    if (!path) return hfail(VXG_ERR_ARG, "null path");
    FILE* f = fopen(path, "rb");
    if (!f) return hfail(VXG_ERR_ARG, "cannot open %s", path);
    std::string text; char buf[65536]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    fclose(f);
    return load_text(text, variant, ignore_collisions, out);
}

extern "C" void vxh_free(vxh_model* m) { delete m; }
extern "C" const vxg_robot_desc* vxh_desc(const vxh_model* m) { return m ? &m->desc : nullptr; }
extern "C" int vxh_get_info(const vxh_model* m, vxh_info* out) {
    if (!m || !out) return hfail(VXG_ERR_ARG, "null argument");
    *out = m->info; return 0;
}
extern "C" int vxh_cell_index(const vxh_model* m, int32_t* out) {
    if (!m || !out) return hfail(VXG_ERR_ARG, "null argument");
    memcpy(out, m->cell_index.data(), m->cell_index.size() * sizeof(int32_t)); return 0;
}

// The CLI loop `while (!StopConditionMet()) TimeStep()` (voxelyzeMain/main.cpp:89-110) replayed on the
// fp64 clock alone: StopConditionMet VX_Sim.cpp:1398-1425, IniCM latch VX_Sim.cpp:1064, EndOfLifetimePosteriorY
// latch :1078-1082, CurTime += dt :1929.
extern "C" int vxh_step_plan_ex(const vxh_model* m, int64_t* total_steps, int64_t* latch_step, int64_t* eol_step, double* final_time) {
    if (!m) return hfail(VXG_ERR_ARG, "null model");
    const vxh_info& I = m->info;
    const double dt = m->desc.sim.dt;
    if (!(dt > 0)) return hfail(VXG_ERR_ARG, "dt is not positive");
    const int type = I.stop_condition_type;
    if (type != 1 && type != 2 && type != 3)
        return hfail(VXG_UNSUPPORTED_FEATURE, "stop condition type %d depends on simulation state; only MAX_TIME_STEPS / MAX_SIM_TIME / TEMP_CYCLES are supported", type);
    double limit = type == 2 ? I.stop_condition_value + I.afterlife_time : (type == 3 ? m->temp_period * I.stop_condition_value : 0);
    if (type != 1 && !(limit / dt < 4e9)) return hfail(VXG_ERR_ARG, "stop condition needs more than 4e9 steps");
    const bool lw = m->variant == VXH_VARIANT_LAND_WATER;
    double t = 0; int64_t step = 0, latch = -1, eol = -1;
    for (;;) {
        bool met = false;
        if (lw || t > I.init_cm_time) {   // the land/water fork has no InitCmTime guard (LW/VX_Sim.cpp StopConditionMet)
            if (type == 1) met = step > (int64_t)(int)(I.stop_condition_value + 0.5);
            else met = t > limit;
        }
        if (met) break;
        if (latch < 0 && t > I.init_cm_time) latch = step;
        if (!lw && eol < 0 && t >= I.stop_condition_value) eol = step;
        t += dt; step++;
        if (step > 4000000000LL) return hfail(VXG_ERR_ARG, "stop condition never met");
    }
    if (total_steps) *total_steps = step;
    if (latch_step) *latch_step = latch;
    if (eol_step) *eol_step = eol;
    if (final_time) *final_time = t;
    return 0;
}

extern "C" int vxh_step_plan(const vxh_model* m, int64_t* total_steps, int64_t* latch_step, double* final_time) {
    return vxh_step_plan_ex(m, total_steps, latch_step, nullptr, final_time);
}

// CVX_SimGA::WriteResultFile (VX_SimGA.cpp:33-168; LW/VX_SimGA.cpp:33-75).  Numbers go through an
// ostringstream exactly like CXML_Rip::Element (Utils/XML_Rip.h:57-59), the layout is tinyxml's SaveFile.
extern "C" int vxh_format_result(const vxh_model* m, const vxg_robot_stats* st, char* buf, int cap, const char* path) {
    if (!m || !st) return hfail(VXG_ERR_ARG, "null argument");
    const double lat = m->info.lattice_dim;
    std::ostringstream o;
    auto el = [&o](const char* tag, double v) { std::ostringstream t; t << v; o << "        <" << tag << ">" << t.str() << "</" << tag << ">\n"; };
    auto elf = [&o](const char* tag, float v) { std::ostringstream t; t << v; o << "        <" << tag << ">" << t.str() << "</" << tag << ">\n"; };
    auto eli = [&o](const char* tag, long long v) { o << "        <" << tag << ">" << v << "</" << tag << ">\n"; };
    o << "<?xml version=\"1.0\" ?>\n<Voxelyze_Sim_Result Version=\"1.0\">\n    <Fitness>\n";
    const double dx = st->com[0] - st->ini_com[0], dy = st->com[1] - st->ini_com[1], dz = st->com[2] - st->ini_com[2];
    if (m->variant == VXH_VARIANT_LAND_WATER) {
        const double nx = dx / lat, ny = dy / lat, nz = dz / lat;
        eli("VoxelNumber", m->info.n_vox);
        elf("normAbsoluteDisplacement", (float)sqrt(nx * nx + ny * ny + nz * nz));
        elf("normDistX", (float)nx); elf("normDistY", (float)ny); elf("normDistZ", (float)nz);
        // shape descriptors are only computed with --computeShapeDescriptors (needs qhull); the stock run leaves -1
        // RobotVolumeStart: the closed surface mesh of the undeformed robot encloses exactly its voxels
        el("RobotVolumeStart", m->info.n_vox * lat * lat * lat); el("ConvexHullVolumeStart", -1);
        el("RobotVolumeEnd", st->mesh_volume);   // -1 from a batch without mesh capacity, like the reference without its mesh
        el("ConvexHullVolumeEnd", -1);
        el("ShapeComplexityStart", -1); el("ShapeComplexityEnd", -1);
    } else {
        const double final_dist = pow(pow(dx, 2) + pow(dy, 2), 0.5) / lat;
        const double post_dist = st->posterior_dist / lat, ant_dist = st->anterior_dist / lat;
        double ant_y = st->anterior_y / lat; if (!(ant_y > 0.0)) ant_y = 0.0;            // getAnteriorY starts from 0
        double post_y = st->posterior_y / lat; if (!(post_y < 100000.0)) post_y = 100000.0;   // getPosteriorY starts from 1e5
        const double eol = st->eol_posterior_y / lat;
        el("NormFinalDist", final_dist);                 // normFinalDist - normFrozenDist, no volume normalisation in scope
        el("NormRegimeDist", post_dist - eol);
        el("NormFrozenDist", 0);
        el("FinalDist", final_dist);
        el("finalDistY", dy / lat);
        el("AnteriorDist", ant_dist);
        el("PosteriorDist", post_dist);
        el("AnteriorY", ant_y);
        el("PosteriorY", post_y);
        el("EndOfLifePosteriorY", eol);
        el("FallAdjPostY", eol);
        eli("NumNonFeetTouchingFloor", st->n_nonfeet_touching_floor);
        eli("NumTouchingFloor", st->n_touching_floor);
        el("Lifetime", st->time - m->info.afterlife_time);
        eli("FoundNeedleInHaystack", 0);
        el("PushDist", 0);
    }
    o << "    </Fitness>\n</Voxelyze_Sim_Result>\n";
    const std::string text = o.str();
    if (buf && cap > 0) { snprintf(buf, (size_t)cap, "%s", text.c_str()); }
    if (path && *path) {
        FILE* f = fopen(path, "wb");
        if (!f) return hfail(VXG_ERR_ARG, "cannot write %s", path);
        fwrite(text.data(), 1, text.size(), f);
        fclose(f);
    }
    return (int)text.size();
}
