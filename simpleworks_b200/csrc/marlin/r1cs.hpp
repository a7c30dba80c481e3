// R1CS instances as the protocol layer sees them: what ark_relations::r1cs::ConstraintSystem
// exposes through to_matrices() plus the two assignment vectors: columns are
// instance variables first (index 0 = the constant one), then witnesses.  The reference's gadgets
// (src/gadgets, src/merkle_tree, src/schnorr_signature) produce such systems on the Rust side; here
// they arrive through the C ABI (swb_r1cs_*) or from the built-in generators below.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

#include "poly.hpp"

namespace swb {
namespace marlin {

struct SparseRow {
    std::vector<std::pair<Fr, uint32_t>> e;   // (coefficient, column)
};

struct R1cs {
    size_t num_instance = 1;                  // includes the constant one
    size_t num_witness = 0;
    std::vector<SparseRow> a, b, c;           // one row per constraint
    std::vector<Fr> instance;                 // [1, x_1, ...]
    std::vector<Fr> witness;
    bool has_assignment = false;

    size_t num_constraints() const { return a.size(); }
    size_t num_variables() const { return num_instance + num_witness; }

    // pad_input_for_indexer_and_prover: instance count up to a power of two (zero inputs);
    // witness columns shift accordingly
    void pad_instance() {
        size_t target = 1;
        while (target < num_instance) target <<= 1;
        const size_t add = target - num_instance;
        if (!add) return;
        auto shift = [&](std::vector<SparseRow>& m) {
            for (auto& row : m)
                for (auto& e : row.e)
                    if (e.second >= num_instance) e.second += (uint32_t)add;
        };
        shift(a); shift(b); shift(c);
        if (has_assignment) instance.resize(target, Fr::zero());
        num_instance = target;
    }
    // make_matrices_square: dummy constraints or dummy witnesses (value one)
    void make_square() {
        const size_t nv = num_variables(), nc = num_constraints();
        if (nc < nv) {
            a.resize(nv); b.resize(nv); c.resize(nv);
        } else if (nv < nc) {
            const size_t add = nc - nv;
            num_witness += add;
            if (has_assignment) witness.resize(witness.size() + add, Fr::one());
        }
    }
    Fr value(uint32_t col) const { return col < num_instance ? instance[col] : witness[col - num_instance]; }
    bool is_satisfied() const {
        if (!has_assignment) return false;
        auto dot = [&](const SparseRow& r) {
            Fr s = Fr::zero();
            for (auto& e : r.e) s = s + e.first * value(e.second);
            return s;
        };
        for (size_t i = 0; i < a.size(); i++)
            if (!(dot(a[i]) * dot(b[i]) == dot(c[i]))) return false;
        return true;
    }
};

// ---- interchange format -------------------------------------------------------------
// What a Rust-side exporter writes from ConstraintSystemRef::to_matrices() + the assignments, so that
// circuits synthesised by the reference's gadgets prove here unchanged (INTEGRATION.md has the
// exporter).  All integers little endian:
//   "SWBR1CS1" | u64 num_instance | u64 num_witness | u64 num_constraints
//   for M in A, B, C: for each constraint: u64 nnz, then nnz x (Fr canonical 32 B, u64 column)
//   u8 has_assignment | [num_instance x Fr canonical | num_witness x Fr canonical]
inline void r1cs_write(const R1cs& cs, std::vector<uint8_t>* out);
inline bool r1cs_read(const uint8_t* p, size_t len, R1cs* cs);

// examples/manual-constraints.rs:16-31: instance [1, a], witness [b], (a - b) * 1 = 0
inline R1cs circuit_manual_constraints(uint64_t a_val, uint64_t b_val) {
    R1cs cs;
    cs.num_instance = 2;
    cs.num_witness = 1;
    cs.a.resize(1); cs.b.resize(1); cs.c.resize(1);
    cs.a[0].e = {{Fr::one(), 1}, {Fr::one().neg(), 2}};
    cs.b[0].e = {{Fr::one(), 0}};
    cs.instance = {Fr::one(), fr_from_u64(a_val)};
    cs.witness = {fr_from_u64(b_val)};
    cs.has_assignment = true;
    return cs;
}
// examples/test-circuit.rs:13-26: two UInt8 witnesses (16 booleans), 8 bitwise equalities
inline R1cs circuit_uint8_equality(uint8_t a_val, uint8_t b_val) {
    R1cs cs;
    cs.num_instance = 1;
    cs.num_witness = 16;
    auto add = [&](SparseRow ra, SparseRow rb, SparseRow rc) { cs.a.push_back(ra); cs.b.push_back(rb); cs.c.push_back(rc); };
    const Fr one = Fr::one(), mone = Fr::one().neg();
    for (uint32_t i = 0; i < 16; i++) {            // booleanity (1 - x) * x = 0
        SparseRow ra, rb, rc;
        ra.e = {{one, 0}, {mone, 1 + i}};
        rb.e = {{one, 1 + i}};
        add(ra, rb, rc);
    }
    for (uint32_t i = 0; i < 8; i++) {             // (b_i - a_i) * 1 = 0
        SparseRow ra, rb, rc;
        ra.e = {{one, 1 + 8 + i}, {mone, 1 + i}};
        rb.e = {{one, 0}};
        add(ra, rb, rc);
    }
    cs.instance = {Fr::one()};
    for (int i = 0; i < 8; i++) cs.witness.push_back(fr_from_u64((a_val >> i) & 1));
    for (int i = 0; i < 8; i++) cs.witness.push_back(fr_from_u64((b_val >> i) & 1));
    cs.has_assignment = true;
    return cs;
}
// synthetic chain (BASELINE config 4): x_i * x_{i+1} = x_{i+2}; one public input x_0
inline R1cs circuit_mul_chain(size_t num_constraints, uint64_t seed0, uint64_t seed1) {
    R1cs cs;
    cs.num_instance = 2;
    cs.num_witness = num_constraints + 1;
    cs.a.resize(num_constraints); cs.b.resize(num_constraints); cs.c.resize(num_constraints);
    std::vector<Fr> x(num_constraints + 2);
    x[0] = fr_from_u64(seed0);
    x[1] = fr_from_u64(seed1);
    for (size_t i = 0; i < num_constraints; i++) x[i + 2] = x[i] * x[i + 1];
    auto col = [&](size_t i) { return (uint32_t)(i == 0 ? 1 : 1 + i); };   // x_0 is instance 1, x_i witness i-1
    const Fr one = Fr::one();
    for (size_t i = 0; i < num_constraints; i++) {
        cs.a[i].e = {{one, col(i)}};
        cs.b[i].e = {{one, col(i + 1)}};
        cs.c[i].e = {{one, col(i + 2)}};
    }
    cs.instance = {Fr::one(), x[0]};
    cs.witness.assign(x.begin() + 1, x.end());
    cs.has_assignment = true;
    return cs;
}

// synthetic general-shape instance (not in the reference: its real circuits need ark-r1cs-std): three
// public inputs, `terms` random (coefficient, variable) pairs per row of A and of B -- columns may
// repeat inside a row, as they can after gadget synthesis -- and one fresh witness per constraint holding
// (A z)(B z), so the instance is satisfied by construction.  A xorshift generator keeps it deterministic.
inline R1cs circuit_random_sparse(size_t num_constraints, uint64_t terms, uint64_t seed) {
    R1cs cs;
    cs.num_instance = 4;
    cs.num_witness = num_constraints;
    cs.a.resize(num_constraints); cs.b.resize(num_constraints); cs.c.resize(num_constraints);
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    auto next = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    std::vector<Fr> z = {Fr::one(), fr_from_u64(next() >> 8), fr_from_u64(next() >> 8), fr_from_u64(next() >> 8)};
    if (terms == 0) terms = 1;
    for (size_t i = 0; i < num_constraints; i++) {
        Fr dot[2] = {Fr::zero(), Fr::zero()};
        for (int side = 0; side < 2; side++) {
            SparseRow& row = side ? cs.b[i] : cs.a[i];
            for (uint64_t t = 0; t < terms; t++) {
                const uint32_t col = (uint32_t)(next() % z.size());
                Fr coef = fr_from_u64((next() >> 40) + 1);
                if (next() & 1) coef = coef.neg();
                row.e.push_back({coef, col});
                dot[side] = dot[side] + coef * z[col];
            }
        }
        cs.c[i].e = {{Fr::one(), (uint32_t)z.size()}};
        z.push_back(dot[0] * dot[1]);
    }
    cs.instance.assign(z.begin(), z.begin() + 4);
    cs.witness.assign(z.begin() + 4, z.end());
    cs.has_assignment = true;
    return cs;
}

}  // namespace marlin
}  // namespace swb

#include "curve_host.hpp"

namespace swb {
namespace marlin {

inline void r1cs_write(const R1cs& cs, std::vector<uint8_t>* out) {
    const char magic[8] = {'S', 'W', 'B', 'R', '1', 'C', 'S', '1'};
    out->insert(out->end(), magic, magic + 8);
    put_u64(*out, cs.num_instance);
    put_u64(*out, cs.num_witness);
    put_u64(*out, cs.num_constraints());
    const std::vector<SparseRow>* ms[3] = {&cs.a, &cs.b, &cs.c};
    for (int m = 0; m < 3; m++)
        for (auto& row : *ms[m]) {
            put_u64(*out, row.e.size());
            for (auto& e : row.e) {
                put_fr_canonical(*out, e.first);
                put_u64(*out, e.second);
            }
        }
    out->push_back(cs.has_assignment ? 1 : 0);
    if (cs.has_assignment) {
        for (auto& x : cs.instance) put_fr_canonical(*out, x);
        for (auto& x : cs.witness) put_fr_canonical(*out, x);
    }
}
inline bool r1cs_read(const uint8_t* p, size_t len, R1cs* cs) {
    const uint8_t* end = p + len;
    if (len < 8 || memcmp(p, "SWBR1CS1", 8) != 0) return false;
    p += 8;
    uint64_t ni, nw, nc;
    if (!get_u64(p, end, &ni) || !get_u64(p, end, &nw) || !get_u64(p, end, &nc) || ni == 0) return false;
    if (nc > ((uint64_t)1 << 32) || ni + nw > ((uint64_t)1 << 32)) return false;
    if (nc > len / 24) return false;          // every constraint needs at least three 8-byte row lengths
    *cs = R1cs();
    cs->num_instance = (size_t)ni;
    cs->num_witness = (size_t)nw;
    std::vector<SparseRow>* ms[3] = {&cs->a, &cs->b, &cs->c};
    for (int m = 0; m < 3; m++) {
        ms[m]->resize((size_t)nc);
        for (auto& row : *ms[m]) {
            uint64_t k;
            // a row may repeat columns (r1cs_write and gadget synthesis both can), so k is bounded by the bytes
            // that are left, not by the variable count: every entry takes 40 bytes
            if (!get_u64(p, end, &k) || k > (uint64_t)(end - p) / 40) return false;
            row.e.resize((size_t)k);
            for (auto& e : row.e) {
                uint64_t col;
                if (!get_fr_canonical(p, end, &e.first) || !get_u64(p, end, &col) || col >= ni + nw) return false;
                e.second = (uint32_t)col;
            }
        }
    }
    if (p >= end) return false;
    cs->has_assignment = *p++ != 0;
    if (cs->has_assignment) {
        if (ni + nw > (uint64_t)(end - p) / 32) return false;      // 32 bytes per value must still be there
        cs->instance.resize((size_t)ni);
        cs->witness.resize((size_t)nw);
        for (auto& x : cs->instance)
            if (!get_fr_canonical(p, end, &x)) return false;
        for (auto& x : cs->witness)
            if (!get_fr_canonical(p, end, &x)) return false;
    }
    return p == end;
}

}  // namespace marlin
}  // namespace swb
