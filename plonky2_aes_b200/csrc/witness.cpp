// libp2witness.so — evaluates the generator program recorded by the Python CircuitBuilder
// (host side; see include/p2witness.h for what it mirrors in the reference stack).
#include "../../include/p2witness.h"
#include "gl64.cuh"
#include <vector>
#include <string.h>
#include <stdlib.h>

struct p2w_program {
    p2w_program_desc d;
    std::vector<int32_t> ops; std::vector<uint64_t> op_consts;
    std::vector<int32_t> lut_lens, lut_off; std::vector<uint16_t> lut_data;
    std::vector<std::vector<int32_t>> key_to_entry;   // per LUT: 65536 -> entry index or -1
    std::vector<int32_t> wire_slot;
    std::vector<int64_t> fixed_pos; std::vector<uint64_t> fixed_val;
    std::vector<int32_t> lookup_counts, lookup_slots, lookup_padding, lookup_off;
    std::vector<int64_t> mult_pos;
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
    *out = p;
    return 0;
}
extern "C" void p2w_program_destroy(p2w_program* p) { delete p; }

static inline int set_slot(std::vector<uint64_t>& val, std::vector<uint8_t>& has, int32_t s, uint64_t v) {
    if (has[s]) return val[s] == v ? 0 : P2W_E_CONFLICT;
    val[s] = v; has[s] = 1;
    return 0;
}

extern "C" int32_t p2w_generate(const p2w_program* p, const int32_t* in_slots, const uint64_t* in_vals, uint32_t num_inputs,
                                uint64_t* wires) {
    const p2w_program_desc& d = p->d;
    std::vector<uint64_t> val(d.num_slots, 0);
    std::vector<uint8_t> has(d.num_slots, 0);
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
        default: return P2W_E_BADARG;
        }
    }
    const size_t n = (size_t)1 << d.log_n, cells = (size_t)d.num_wires * n;
    const int32_t* ws = p->wire_slot.data();
    for (size_t i = 0; i < cells; i++) wires[i] = ws[i] >= 0 ? val[ws[i]] : 0;
    for (uint32_t i = 0; i < d.num_fixed; i++) wires[p->fixed_pos[i]] = p->fixed_val[i];
    // set_lookup_wires: multiplicities
    for (uint32_t l = 0; l < d.num_luts; l++) {
        std::vector<uint64_t> mult(p->lut_lens[l], 0);
        const int32_t* ls = p->lookup_slots.data() + p->lookup_off[l];
        for (int32_t i = 0; i < p->lookup_counts[l]; i++) {
            if (!has[ls[i]]) return P2W_E_UNSET;
            uint64_t x = val[ls[i]];
            if (x > 0xFFFF || p->key_to_entry[l][x] < 0) return P2W_E_LOOKUP;
            mult[p->key_to_entry[l][x]]++;
        }
        if (p->lut_lens[l]) mult[0] += p->lookup_padding[l];
        const int64_t* mp = p->mult_pos.data() + p->lut_off[l];
        for (int32_t e = 0; e < p->lut_lens[l]; e++) wires[mp[e]] = mult[e];
    }
    return 0;
}

extern "C" int32_t p2w_generate_many(const p2w_program* p, const int32_t* in_slots, const uint64_t* in_vals, uint32_t num_inputs,
                                     uint32_t count, uint64_t* wires) {
    const size_t cells = (size_t)p->d.num_wires << p->d.log_n;
    int32_t rc_all = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (uint32_t w = 0; w < count; w++) {
        int32_t rc = p2w_generate(p, in_slots, in_vals + (size_t)w * num_inputs, num_inputs, wires + w * cells);
        if (rc) {
#pragma omp critical
            rc_all = rc;
        }
    }
    return rc_all;
}
