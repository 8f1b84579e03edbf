// libp2witness.so — evaluates the generator program recorded by the Python CircuitBuilder
// (host side; see include/p2witness.h for what it mirrors in the reference stack).
#include "../../include/p2witness.h"
#include "gl64.cuh"
#include <vector>
#include <string.h>
#include <stdlib.h>
#include <omp.h>

static const uint64_t W_RC[360] = {
#include "poseidon_rc.inc"
};
#include "poseidon_fast.inc"
static inline uint64_t w_sbox(uint64_t x) { uint64_t x2 = gl_mul(x, x), x4 = gl_mul(x2, x2), x3 = gl_mul(x, x2); return gl_mul(x3, x4); }
static inline void w_mds(uint64_t s[12]) {
    static const uint64_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint64_t o[12];
    for (int r = 0; r < 12; r++) {
        unsigned __int128 acc = 0;
        for (int i = 0; i < 12; i++) acc += (unsigned __int128)s[(i + r) % 12] * C[i];
        if (r == 0) acc += (unsigned __int128)s[0] * 8;
        o[r] = gl_canon(gl_reduce128_lazy((uint64_t)acc, (uint64_t)(acc >> 64)));
    }
    memcpy(s, o, sizeof(o));
}
// PoseidonGate generator with swap = 0: fills wires 12..134 of the row (trace[c - 12]) and returns the output state
static void w_poseidon_gate(const uint64_t in[12], uint64_t trace[123]) {
    uint64_t st[12]; memcpy(st, in, sizeof(st));
    memset(trace, 0, 123 * sizeof(uint64_t));           // swap (24) and deltas (25..28) are zero
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], W_RC[12 * r + i]);
        if (r != 0) for (int i = 0; i < 12; i++) trace[29 + 12 * (r - 1) + i - 12] = st[i];
        for (int i = 0; i < 12; i++) st[i] = w_sbox(st[i]);
        w_mds(st);
    }
    for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], PFAST_FIRST_C[i]);
    { uint64_t t[11]; for (int r = 0; r < 11; r++) { uint64_t a = 0; for (int c = 0; c < 11; c++) a = gl_add(a, gl_mul(st[c + 1], PFAST_INIT[r * 11 + c])); t[r] = a; }
      for (int r = 0; r < 11; r++) st[r + 1] = t[r]; }
    for (int r = 0; r < 22; r++) {
        trace[65 + r - 12] = st[0];
        st[0] = gl_add(w_sbox(st[0]), PFAST_K[r]);
        uint64_t s0 = gl_mul(st[0], 25);
        for (int j = 0; j < 11; j++) s0 = gl_add(s0, gl_mul(st[j + 1], PFAST_VROW[r * 11 + j]));
        for (int j = 0; j < 11; j++) st[j + 1] = gl_add(st[j + 1], gl_mul(st[0], PFAST_WCOL[r * 11 + j]));
        st[0] = s0;
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], W_RC[12 * (26 + r) + i]);
        for (int i = 0; i < 12; i++) trace[87 + 12 * r + i - 12] = st[i];
        for (int i = 0; i < 12; i++) st[i] = w_sbox(st[i]);
        w_mds(st);
    }
    for (int i = 0; i < 12; i++) trace[i] = st[i];         // outputs, wires 12..23
}

struct p2w_program {
    p2w_program_desc d;
    std::vector<int32_t> ops; std::vector<uint64_t> op_consts;
    std::vector<int32_t> lut_lens, lut_off; std::vector<uint16_t> lut_data;
    std::vector<std::vector<int32_t>> key_to_entry;   // per LUT: 65536 -> entry index or -1
    std::vector<int32_t> wire_slot;
    std::vector<int64_t> fixed_pos; std::vector<uint64_t> fixed_val;
    std::vector<int32_t> lookup_counts, lookup_slots, lookup_padding, lookup_off;
    std::vector<int64_t> mult_pos;
    std::vector<int32_t> poseidon_rows;
    // extended slot vector ("partition values"): [program slots | multiplicities of every LUT entry |
    // 111 internal wires per PoseidonGate row]; wire_map indexes it (-1 = empty cell)
    uint32_t ext_mult, ext_pos, ext_total;
    std::vector<int32_t> wire_map;
    std::vector<int64_t> fixed_pos_eff; std::vector<uint64_t> fixed_val_eff;   // fixed cells not overridden by the map
};

extern "C" int32_t p2w_program_create(const p2w_program_desc* d, p2w_program** out) {
    if (!d || !out) return P2W_E_BADARG;
    p2w_program* p = new p2w_program();
    p->d = *d;
    p->ops.assign(d->ops, d->ops + (size_t)d->num_ops * 6);
    p->op_consts.assign(d->op_consts, d->op_consts + (size_t)d->num_ops * 2);
    p->lut_lens.assign(d->lut_lens, d->lut_lens + d->num_luts);
    size_t tot = 0, totl = 0;
    for (uint32_t i = 0; i < d->num_luts; i++) { p->lut_off.push_back((int32_t)tot); tot += d->lut_lens[i]; }
    p->lut_data.assign(d->lut_data, d->lut_data + 2 * tot);
    p->key_to_entry.resize(d->num_luts);
    for (uint32_t i = 0; i < d->num_luts; i++) {
        p->key_to_entry[i].assign(65536, -1);
        for (int32_t e = d->lut_lens[i] - 1; e >= 0; e--)   // first occurrence wins, as a HashMap built in order would keep the last; tables have unique keys
            p->key_to_entry[i][p->lut_data[2 * ((size_t)p->lut_off[i] + e)]] = e;
    }
    size_t cells = (size_t)d->num_wires << d->log_n;
    p->wire_slot.assign(d->wire_slot, d->wire_slot + cells);
    p->fixed_pos.assign(d->fixed_pos, d->fixed_pos + d->num_fixed);
    p->fixed_val.assign(d->fixed_val, d->fixed_val + d->num_fixed);
    p->lookup_counts.assign(d->lookup_counts, d->lookup_counts + d->num_luts);
    p->lookup_padding.assign(d->lookup_padding, d->lookup_padding + d->num_luts);
    for (uint32_t i = 0; i < d->num_luts; i++) { p->lookup_off.push_back((int32_t)totl); totl += d->lookup_counts[i]; }
    p->lookup_slots.assign(d->lookup_slots, d->lookup_slots + totl);
    p->mult_pos.assign(d->mult_pos, d->mult_pos + tot);
    if (d->num_poseidon) p->poseidon_rows.assign(d->poseidon_rows, d->poseidon_rows + (size_t)25 * d->num_poseidon);
    // wire map over the extended slot vector
    const size_t n = (size_t)1 << d->log_n;
    p->ext_mult = d->num_slots;
    p->ext_pos = p->ext_mult + (uint32_t)tot;
    p->ext_total = p->ext_pos + 111u * d->num_poseidon;
    p->wire_map = p->wire_slot;
    for (uint32_t k = 0; k < d->num_poseidon; k++) {
        const size_t row = (size_t)p->poseidon_rows[(size_t)25 * k];
        for (int c = 24; c < 135; c++) p->wire_map[(size_t)c * n + row] = (int32_t)(p->ext_pos + 111u * k + (uint32_t)(c - 24));
    }
    for (size_t e = 0; e < tot; e++) p->wire_map[p->mult_pos[e]] = (int32_t)(p->ext_mult + e);
    for (uint32_t i = 0; i < d->num_fixed; i++) {
        const int32_t m = p->wire_map[p->fixed_pos[i]];
        if (m >= (int32_t)p->ext_mult) continue;            // a multiplicity / Poseidon cell wins, as in the fill order
        p->wire_map[p->fixed_pos[i]] = -1;
        p->fixed_pos_eff.push_back(p->fixed_pos[i]); p->fixed_val_eff.push_back(p->fixed_val[i]);
    }
    *out = p;
    return 0;
}
extern "C" void p2w_program_destroy(p2w_program* p) { delete p; }

static inline int set_slot(std::vector<uint64_t>& val, std::vector<uint8_t>& has, int32_t s, uint64_t v) {
    if (has[s]) return val[s] == v ? 0 : P2W_E_CONFLICT;
    val[s] = v; has[s] = 1;
    return 0;
}

// Runs the generators and set_lookup_wires: fills the extended slot vector (p->ext_total words).
extern "C" int32_t p2w_generate_slots(const p2w_program* p, const int32_t* in_slots, const uint64_t* in_vals, uint32_t num_inputs,
                                      uint64_t* ext) {
    if (!p || !ext || (num_inputs && (!in_slots || !in_vals))) return P2W_E_BADARG;
    const p2w_program_desc& d = p->d;
    std::vector<uint64_t> val(d.num_slots, 0);
    std::vector<uint8_t> has(d.num_slots, 0);
    std::vector<uint64_t> ptrace((size_t)123 * d.num_poseidon);
    int rc;
    for (uint32_t i = 0; i < num_inputs; i++) {
        if (in_slots[i] < 0 || (uint32_t)in_slots[i] >= d.num_slots || in_vals[i] >= GL_P) return P2W_E_BADARG;
        if ((rc = set_slot(val, has, in_slots[i], in_vals[i]))) return rc;
    }
    const int32_t* op = p->ops.data();
    const uint64_t* oc = p->op_consts.data();
    for (uint32_t k = 0; k < d.num_ops; k++, op += 6, oc += 2) {
        switch (op[0]) {
        case P2W_OP_ARITH: {
            if (!has[op[2]] || !has[op[3]] || !has[op[4]]) return P2W_E_UNSET;
            uint64_t r = gl_add(gl_mul(oc[0], gl_mul(val[op[2]], val[op[3]])), gl_mul(oc[1], val[op[4]]));
            if ((rc = set_slot(val, has, op[1], r))) return rc;
            break; }
        case P2W_OP_LOOKUP: {
            if (!has[op[2]]) return P2W_E_UNSET;
            uint64_t x = val[op[2]];
            int32_t lut = op[5];
            if (x > 0xFFFF) return P2W_E_LOOKUP;
            int32_t e = p->key_to_entry[lut][x];
            if (e < 0) return P2W_E_LOOKUP;
            if ((rc = set_slot(val, has, op[1], p->lut_data[2 * ((size_t)p->lut_off[lut] + e) + 1]))) return rc;
            break; }
        case P2W_OP_EQ: {
            if (!has[op[3]] || !has[op[4]]) return P2W_E_UNSET;
            uint64_t diff = gl_sub(val[op[3]], val[op[4]]);
            if ((rc = set_slot(val, has, op[1], diff == 0 ? 1 : 0))) return rc;
            if ((rc = set_slot(val, has, op[2], diff == 0 ? 0 : gl_inv(diff)))) return rc;
            break; }
        case P2W_OP_CONST:
            if ((rc = set_slot(val, has, op[1], oc[0]))) return rc;
            break;
        case P2W_OP_POSEIDON: {
            if (op[1] < 0 || (uint32_t)op[1] >= d.num_poseidon) return P2W_E_BADARG;
            const int32_t* pr = p->poseidon_rows.data() + (size_t)25 * op[1];
            uint64_t in[12];
            for (int i = 0; i < 12; i++) { if (!has[pr[1 + i]]) return P2W_E_UNSET; in[i] = val[pr[1 + i]]; }
            uint64_t* tr = ptrace.data() + (size_t)123 * op[1];
            w_poseidon_gate(in, tr);
            for (int i = 0; i < 12; i++) if ((rc = set_slot(val, has, pr[13 + i], tr[i]))) return rc;
            break; }
        default: return P2W_E_BADARG;
        }
    }
    memcpy(ext, val.data(), (size_t)d.num_slots * sizeof(uint64_t));
    for (uint32_t k = 0; k < d.num_poseidon; k++)           // internal wires 24..134 of every PoseidonGate row
        memcpy(ext + p->ext_pos + (size_t)111 * k, ptrace.data() + (size_t)123 * k + 12, 111 * sizeof(uint64_t));
    // set_lookup_wires: multiplicities
    for (uint32_t l = 0; l < d.num_luts; l++) {
        uint64_t* mult = ext + p->ext_mult + p->lut_off[l];
        memset(mult, 0, (size_t)p->lut_lens[l] * sizeof(uint64_t));
        const int32_t* ls = p->lookup_slots.data() + p->lookup_off[l];
        for (int32_t i = 0; i < p->lookup_counts[l]; i++) {
            if (!has[ls[i]]) return P2W_E_UNSET;
            uint64_t x = val[ls[i]];
            if (x > 0xFFFF || p->key_to_entry[l][x] < 0) return P2W_E_LOOKUP;
            mult[p->key_to_entry[l][x]]++;
        }
        if (p->lut_lens[l]) mult[0] += p->lookup_padding[l];
    }
    return 0;
}
extern "C" uint32_t p2w_ext_slots(const p2w_program* p) { return p ? p->ext_total : 0; }
extern "C" int32_t p2w_wire_map(const p2w_program* p, int32_t* map_out) {
    if (!p || !map_out) return P2W_E_BADARG;
    memcpy(map_out, p->wire_map.data(), p->wire_map.size() * sizeof(int32_t));
    return 0;
}
extern "C" int32_t p2w_fixed_cells(const p2w_program* p, uint32_t* count, const int64_t** pos, const uint64_t** val) {
    if (!p || !count || !pos || !val) return P2W_E_BADARG;
    *count = (uint32_t)p->fixed_pos_eff.size(); *pos = p->fixed_pos_eff.data(); *val = p->fixed_val_eff.data();
    return 0;
}
// PartitionWitness::full_witness on the host: wires[cell] = ext[wire_map[cell]], then the fixed cells.
// (p2g_prove_slots does the same gather on the device.)
extern "C" int32_t p2w_generate(const p2w_program* p, const int32_t* in_slots, const uint64_t* in_vals, uint32_t num_inputs,
                                uint64_t* wires) {
    if (!p || !wires) return P2W_E_BADARG;
    std::vector<uint64_t> ext(p->ext_total);
    int32_t rc = p2w_generate_slots(p, in_slots, in_vals, num_inputs, ext.data());
    if (rc) return rc;
    const size_t cells = (size_t)p->d.num_wires << p->d.log_n;
    const int32_t* wm = p->wire_map.data();
    for (size_t i = 0; i < cells; i++) wires[i] = wm[i] >= 0 ? ext[wm[i]] : 0;
    for (size_t i = 0; i < p->fixed_pos_eff.size(); i++) wires[p->fixed_pos_eff[i]] = p->fixed_val_eff[i];
    return 0;
}

// OpenMP threads used by the *_many entry points (launchers such as torchrun export OMP_NUM_THREADS=1)
static int g_p2w_threads = 0;
extern "C" void p2w_set_num_threads(int32_t n) { g_p2w_threads = n > 0 ? n : 0; }
extern "C" int32_t p2w_num_threads(void) { return g_p2w_threads > 0 ? g_p2w_threads : omp_get_max_threads(); }

extern "C" int32_t p2w_generate_slots_many(const p2w_program* p, const int32_t* in_slots, const uint64_t* in_vals, uint32_t num_inputs,
                                           uint32_t count, uint64_t* ext_out) {
    if (!p || !ext_out) return P2W_E_BADARG;
    int32_t rc_all = 0;
#pragma omp parallel for schedule(dynamic) num_threads(p2w_num_threads())
    for (int64_t w = 0; w < (int64_t)count; w++) {
        int32_t rc = p2w_generate_slots(p, in_slots, in_vals + (size_t)w * num_inputs, num_inputs, ext_out + (size_t)w * p->ext_total);
        if (rc) {
#pragma omp critical
            rc_all = rc;
        }
    }
    return rc_all;
}

extern "C" int32_t p2w_generate_many(const p2w_program* p, const int32_t* in_slots, const uint64_t* in_vals, uint32_t num_inputs,
                                     uint32_t count, uint64_t* wires) {
    const size_t cells = (size_t)p->d.num_wires << p->d.log_n;
    int32_t rc_all = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(p2w_num_threads())
    for (uint32_t w = 0; w < count; w++) {
        int32_t rc = p2w_generate(p, in_slots, in_vals + (size_t)w * num_inputs, num_inputs, wires + w * cells);
        if (rc) {
#pragma omp critical
            rc_all = rc;
        }
    }
    return rc_all;
}
