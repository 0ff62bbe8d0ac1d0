// Splitting-string compiler (host only, no CUDA calls).
//
// Implements the behaviour of LangevinSplittingIntegrator.__init__ and
// ContinuousLangevinSplittingIntegrator.__init__
// (reference: benchmark/integrators/langevin.py:89-121,189-196 and :277-302,361-364,406-408):
// same tokenisation, same substep lengths, same set of accepted / rejected strings.
#include <cctype>
#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <sstream>

#include "lsi_internal.h"

static thread_local std::string g_last_error;

int lsi_set_error(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

LSI_EXPORT const char* lsi_last_error(void) { return g_last_error.c_str(); }
LSI_EXPORT int lsi_abi_version(void) { return LSI_ABI_VERSION; }

namespace {

struct Token {
    char head;          // 'R', 'V', 'O' or other
    std::string label;  // text after the first character ("" for bare tokens)
    std::string text;
};

std::vector<Token> tokenize(const char* s)
{
    std::vector<Token> out;
    std::istringstream in(s);  // str.split(): any whitespace run separates tokens
    std::string w;
    while (in >> w) {
        for (auto& c : w) c = (char)std::toupper((unsigned char)c);
        out.push_back(Token{w[0], w.substr(1), w});
    }
    return out;
}

// "V<i>" with 0 <= i < 32 written without sign or leading zeros, as "V{}".format(i) prints it
bool is_group_token(const Token& t)
{
    if (t.head != 'V' || t.label.empty() || t.label.size() > 2) return false;
    for (char c : t.label)
        if (!std::isdigit((unsigned char)c)) return false;
    if (t.label.size() == 2 && t.label[0] == '0') return false;
    return std::stoi(t.label) < 32;
}

bool parse_group(const std::string& label, int* g)
{
    if (label.empty()) { *g = -1; return true; }
    for (char c : label)
        if (!std::isdigit((unsigned char)c)) return false;
    if (label.size() > 6) return false;
    *g = std::stoi(label);
    return true;
}

}  // namespace

LSI_EXPORT int lsi_compile(const char* splitting, double dt, double gamma, int measure_shadow_work,
                           int measure_heat, int override_checks, lsi_program** out)
{
    if (!splitting || !out) return lsi_set_error(LSI_ERR_ARG, "lsi_compile: NULL argument");
    const std::vector<Token> toks = tokenize(splitting);

    int n_R = 0, n_V = 0, n_O = 0;
    std::set<std::string> v_tokens;
    bool any_V = false;
    for (const Token& t : toks) {
        n_R += (t.text == "R");
        n_V += (t.text == "V");
        n_O += (t.text == "O");
        if (t.head == 'V') { v_tokens.insert(t.text); any_V = true; }
    }
    const bool mts = v_tokens.size() > 1;
    // n_Vs[label] counts every token whose tail equals the label (langevin.py:103)
    std::map<std::string, int> n_Vs;
    if (mts)
        for (const std::string& vt : v_tokens) {
            const std::string label = vt.substr(1);
            int n = 0;
            for (const Token& t : toks) n += (t.label == label);
            n_Vs[label] = n;
        }

    if (!override_checks) {
        if (n_R == 0) return lsi_set_error(LSI_ERR_ASSERT, "splitting '%s' contains no R step", splitting);
        if (!any_V) return lsi_set_error(LSI_ERR_ASSERT, "splitting '%s' contains no V step", splitting);
        if (n_O == 0) return lsi_set_error(LSI_ERR_ASSERT, "splitting '%s' contains no O step", splitting);
        for (const Token& t : toks) {
            const bool ok = t.text == "R" || t.text == "V" || t.text == "O" || is_group_token(t);
            if (!ok) return lsi_set_error(LSI_ERR_ASSERT, "invalid token '%s' in splitting '%s'", t.text.c_str(), splitting);
        }
        if (mts && n_V > 0)
            return lsi_set_error(LSI_ERR_VALUE, "Splitting string includes an evaluation of all forces and "
                                                "evaluation of subsets of forces.");
    }

    const double h = dt / (double)(n_O > 1 ? n_O : 1);
    const double a = std::exp(-gamma * h);
    const double b = std::sqrt(1.0 - std::exp(-2.0 * gamma * h));

    lsi_program* p = new lsi_program();
    p->dt = dt; p->gamma = gamma; p->mts = mts; p->n_O = 0;
    p->measure_shadow_work = measure_shadow_work != 0;
    p->measure_heat = measure_heat != 0;
    for (const Token& t : toks) {
        if (t.text == "O") {
            p->ops.push_back(lsi_op{LSI_OP_O, -1, 1.0 / (double)(n_O > 1 ? n_O : 1), a, b});
            p->n_O++;
        } else if (t.text == "R") {
            p->ops.push_back(lsi_op{LSI_OP_R, -1, 1.0 / (double)n_R, 0.0, 0.0});
        } else if (t.head == 'V') {
            int g = -1;
            if (!parse_group(t.label, &g)) {
                delete p;
                return lsi_set_error(LSI_ERR_VALUE, "cannot parse force group of token '%s'", t.text.c_str());
            }
            int denom = mts ? n_Vs[t.label] : n_V;
            if (denom == 0) {
                delete p;
                return lsi_set_error(LSI_ERR_VALUE, "splitting '%s': kick length dt / 0 (a lone force-group token is "
                                                    "not multi-timestep; use bare V)", splitting);
            }
            // non-MTS strings use all forces whatever the label says (langevin.py:157)
            p->ops.push_back(lsi_op{LSI_OP_V, mts ? g : -1, 1.0 / (double)denom, 0.0, 0.0});
        }
        // any other token emits nothing (langevin.py:177-183), reachable only with override_checks
    }
    *out = p;
    return LSI_OK;
}

LSI_EXPORT int lsi_compile_continuous(const char* const* tokens, const double* fractions, int n, double dt,
                                      double gamma, int measure_shadow_work, int measure_heat,
                                      int override_checks, lsi_program** out)
{
    if (!tokens || !fractions || !out || n < 0) return lsi_set_error(LSI_ERR_ARG, "lsi_compile_continuous: bad argument");
    std::vector<Token> toks;
    for (int i = 0; i < n; ++i) {
        if (!tokens[i] || !tokens[i][0]) return lsi_set_error(LSI_ERR_ARG, "empty token %d", i);
        std::string w(tokens[i]);  // the continuous variant does not upper-case (langevin.py:371-377)
        toks.push_back(Token{w[0], w.substr(1), w});
    }
    if (!override_checks) {
        bool has_R = false, has_V = false, has_O = false;
        std::map<std::string, double> lengths;
        for (int i = 0; i < n; ++i) {
            has_R |= toks[i].text == "R";
            has_V |= toks[i].head == 'V';
            has_O |= toks[i].text == "O";
            lengths[toks[i].text] += fractions[i];
        }
        if (!has_R || !has_V || !has_O)
            return lsi_set_error(LSI_ERR_ASSERT, "continuous splitting needs R, V and O steps");
        // lengths["V"] is looked up unconditionally (langevin.py:300-302): a missing bare V is a KeyError there
        if (!lengths.count("V")) return lsi_set_error(LSI_ERR_KEY, "continuous splitting has no bare 'V' entry");
        if (lengths["R"] != 1.0 || lengths["O"] != 1.0 || lengths["V"] != 1.0)
            return lsi_set_error(LSI_ERR_ASSERT, "substep fractions of R, O and V must each sum to 1");
    }
    lsi_program* p = new lsi_program();
    p->dt = dt; p->gamma = gamma; p->mts = false; p->n_O = 0;
    p->measure_shadow_work = measure_shadow_work != 0;
    p->measure_heat = measure_heat != 0;
    std::set<std::string> v_tokens;
    for (int i = 0; i < n; ++i) {
        if (!(fractions[i] > 0.0)) continue;
        const Token& t = toks[i];
        if (t.text == "O") {
            const double h = dt * fractions[i];
            p->ops.push_back(lsi_op{LSI_OP_O, -1, fractions[i], std::exp(-gamma * h), std::sqrt(1.0 - std::exp(-2.0 * gamma * h))});
            p->n_O++;
        } else if (t.text == "R") {
            p->ops.push_back(lsi_op{LSI_OP_R, -1, fractions[i], 0.0, 0.0});
        } else if (t.head == 'V') {
            int g = -1;
            if (!parse_group(t.label, &g)) {
                delete p;
                return lsi_set_error(LSI_ERR_VALUE, "cannot parse force group of token '%s'", t.text.c_str());
            }
            v_tokens.insert(t.text);
            p->ops.push_back(lsi_op{LSI_OP_V, g, fractions[i], 0.0, 0.0});
        }
    }
    p->mts = v_tokens.size() > 1;
    *out = p;
    return LSI_OK;
}

LSI_EXPORT int lsi_program_num_ops(const lsi_program* p) { return p ? (int)p->ops.size() : LSI_ERR_ARG; }
LSI_EXPORT int lsi_program_is_mts(const lsi_program* p) { return p ? (int)p->mts : LSI_ERR_ARG; }
LSI_EXPORT int lsi_program_num_O(const lsi_program* p) { return p ? p->n_O : LSI_ERR_ARG; }

LSI_EXPORT int lsi_program_get_op(const lsi_program* p, int i, int* kind, int* group, double* fraction, double* a, double* b)
{
    if (!p || i < 0 || i >= (int)p->ops.size()) return lsi_set_error(LSI_ERR_ARG, "lsi_program_get_op: bad index %d", i);
    const lsi_op& op = p->ops[i];
    if (kind) *kind = op.kind;
    if (group) *group = op.group;
    if (fraction) *fraction = op.fraction;
    if (a) *a = op.a;
    if (b) *b = op.b;
    return LSI_OK;
}

LSI_EXPORT double lsi_program_get_step_size(const lsi_program* p) { return p ? p->dt : NAN; }

LSI_EXPORT int lsi_program_set_step_size(lsi_program* p, double dt)
{
    if (!p || !(dt == dt)) return lsi_set_error(LSI_ERR_ARG, "lsi_program_set_step_size: bad argument");
    p->dt = dt;  // a, b deliberately untouched (reference behaviour)
    return LSI_OK;
}

LSI_EXPORT int lsi_program_flags(const lsi_program* p, int* msw, int* mh)
{
    if (!p) return lsi_set_error(LSI_ERR_ARG, "NULL program");
    if (msw) *msw = p->measure_shadow_work;
    if (mh) *mh = p->measure_heat;
    return LSI_OK;
}

LSI_EXPORT void lsi_program_destroy(lsi_program* p)
{
    if (!p) return;
    if (!p->staged.empty()) lsi_program_free_staged(p);
    delete p;
}
