"""Host-side mirror of plonky2's `CircuitBuilder` / `CircuitData` / `PartialWitness`, reduced
to what the 0xPARC/plonky2-aes AES and AES-GCM gadgets use (virtual targets, constants,
copy constraints, base arithmetic ops, `is_equal`/`select`, lookup tables) — the interface the
reference drives at e.g. /root/reference/aes-gcm/src/circuit_gcm.rs:49-172 and
/root/reference/aes-gcm/src/circuit_aes.rs:176-358.

Circuit building and witness generation are host work in the reference too (north_star); this
module is the workload generator and the caller of the GPU hot path: `CircuitData.prove`
hands the full wire matrix to `p2g_prove` (libp2gpu.so).  Gate placement follows upstream's
rules (plonk/circuit_builder.rs: `find_slot`, `add_all_lookups`, selector grouping in
gates/selectors.rs, sigma construction in plonk/permutation_argument.rs) as recalled; the exact
row order of the Rust builder cannot be checked here (parity unpinned).
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import ffi
from .field import gl_mul, powers as gl_powers

P = 0xFFFFFFFF00000001
NUM_WIRES = 135

GATE_NOOP, GATE_CONSTANT, GATE_PUBLIC_INPUT, GATE_ARITHMETIC, GATE_LOOKUP, GATE_LOOKUP_TABLE, GATE_POSEIDON = range(7)
# (degree, id string) — the sort key upstream uses in CircuitBuilder::build
GATE_META = {
    GATE_NOOP: (0, "NoopGate"),
    GATE_LOOKUP: (0, "LookupGate"),
    GATE_LOOKUP_TABLE: (0, "LookupTableGate"),
    GATE_CONSTANT: (1, "ConstantGate { num_consts: 2 }"),
    GATE_PUBLIC_INPUT: (1, "PublicInputGate"),
    GATE_ARITHMETIC: (3, "ArithmeticGate { num_ops: 20 }"),
    GATE_POSEIDON: (7, "PoseidonGate(PhantomData<GoldilocksField>)<WIDTH=12>"),
}
UNUSED_SELECTOR = (1 << 32) - 1
OP_ARITH, OP_LOOKUP, OP_EQ, OP_CONST, OP_POSEIDON = 0, 1, 2, 3, 4


@dataclass
class FriConfig:
    rate_bits: int = 3
    cap_height: int = 4
    proof_of_work_bits: int = 16
    reduction_arity_bits: int = 4      # ConstantArityBits(4, 5)
    final_poly_bits: int = 5
    num_query_rounds: int = 28


@dataclass
class CircuitConfig:
    num_wires: int = NUM_WIRES
    num_routed_wires: int = 80
    num_constants: int = 2
    security_bits: int = 100
    num_challenges: int = 2
    zero_knowledge: bool = False
    max_quotient_degree_factor: int = 8
    fri_config: FriConfig = field(default_factory=FriConfig)

    @staticmethod
    def standard_recursion_config():
        """plonky2 CircuitConfig::standard_recursion_config (the config of every reference test,
        e.g. /root/reference/aes-gcm/src/circuit_gcm.rs:757)."""
        return CircuitConfig()

    @staticmethod
    def standard_recursion_zk_config():
        raise NotImplementedError("zero_knowledge=true adds random blinding; north_star excludes it")


@dataclass(frozen=True)
class BoolTarget:
    target: int


def wire(row, col):
    return row * NUM_WIRES + col


class PartialWitness:
    """PartialWitness::new() / set_target: values the caller fixes before proving."""

    def __init__(self):
        self.values = {}

    def set_target(self, target, value):
        value = int(value) % P
        old = self.values.get(target)
        if old is not None and old != value:
            raise ValueError(f"target {target} was set twice with different values")
        self.values[target] = value


class CircuitBuilder:
    def __init__(self, config=None):
        self.config = config or CircuitConfig.standard_recursion_config()
        assert not self.config.zero_knowledge
        self.num_virtual = 0
        self.rows = []             # (gate kind, [constants])
        self.copy_a, self.copy_b = [], []
        self.ops = []              # (kind, s0, s1, s2, s3, s4, c0, c1) on targets
        self.constants_to_targets = {}
        self.arith_slots = {}      # (c0, c1) -> (row, next op index)
        self.arith_cache = {}
        self.luts = []             # list of list[(inp, out)]
        self.lut_to_lookups = []   # per LUT: [(inp target, out target)]
        self.public_inputs = []
        self.poseidon_rows = []    # (row, [12 input wire targets], [12 output wire targets])
        self.ops_per_arith_row = self.config.num_routed_wires // 4

    # ---- targets -----------------------------------------------------------------------
    def add_virtual_target(self):
        self.num_virtual += 1
        return -self.num_virtual

    def add_virtual_targets(self, n):
        return [self.add_virtual_target() for _ in range(n)]

    def register_public_input(self, t):
        """CircuitBuilder::register_public_input: the proof carries the value of `t`."""
        self.public_inputs.append(t)

    def register_public_inputs(self, ts):
        for t in ts:
            self.register_public_input(t)

    def add_virtual_bool_target_unsafe(self):
        return BoolTarget(self.add_virtual_target())

    def constant(self, c):
        c = int(c) % P
        t = self.constants_to_targets.get(c)
        if t is None:
            t = self.add_virtual_target()
            self.constants_to_targets[c] = t
            self._const_of = getattr(self, "_const_of", {})
            self._const_of[t] = c
        return t

    def zero(self):
        return self.constant(0)

    def one(self):
        return self.constant(1)

    def target_as_constant(self, t):
        return getattr(self, "_const_of", {}).get(t)

    def connect(self, a, b):
        self.copy_a.append(a)
        self.copy_b.append(b)

    def assert_zero(self, x):
        self.connect(x, self.zero())

    def num_gates(self):
        return len(self.rows)

    def add_gate(self, kind, constants=()):
        self.rows.append((kind, list(constants)))
        return len(self.rows) - 1

    # ---- base arithmetic (gadgets/arithmetic.rs) -------------------------------------------
    def arithmetic(self, c0, c1, m0, m1, addend):
        """c0 * m0 * m1 + c1 * addend via one ArithmeticGate operation."""
        c0 %= P
        c1 %= P
        k0, k1, ka = self.target_as_constant(m0), self.target_as_constant(m1), self.target_as_constant(addend)
        first_zero = c0 == 0 or k0 == 0 or k1 == 0
        second_zero = c1 == 0 or ka == 0
        if k0 is not None and k1 is not None and ka is not None:
            return self.constant((c0 * k0 * k1 + c1 * ka) % P)
        if first_zero and second_zero:
            return self.zero()
        if first_zero and c1 == 1:
            return addend
        if second_zero and c0 == 1:
            if k0 == 1:
                return m1
            if k1 == 1:
                return m0
        key = (c0, c1, m0, m1, addend)
        hit = self.arith_cache.get(key)
        if hit is not None:
            return hit
        row, i = self.arith_slots.get((c0, c1), (None, 0))
        if row is None:
            row = self.add_gate(GATE_ARITHMETIC, (c0, c1))
            i = 0
        nxt = i + 1
        self.arith_slots[(c0, c1)] = (row, nxt) if nxt < self.ops_per_arith_row else (None, 0)
        w0, w1, w2, w3 = (wire(row, 4 * i + k) for k in range(4))
        self.connect(m0, w0)
        self.connect(m1, w1)
        self.connect(addend, w2)
        self.ops.append((OP_ARITH, w3, w0, w1, w2, 0, c0, c1))
        self.arith_cache[key] = w3
        return w3

    def add(self, a, b):
        return self.arithmetic(1, 1, a, self.one(), b)

    def sub(self, a, b):
        return self.arithmetic(1, P - 1, a, self.one(), b)

    def mul(self, a, b):
        return self.arithmetic(1, 0, a, b, self.zero())

    def mul_add(self, a, b, c):
        return self.arithmetic(1, 1, a, b, c)

    def mul_sub(self, a, b, c):
        return self.arithmetic(1, P - 1, a, b, c)

    def mul_const(self, c, x):
        return self.arithmetic(c, 0, x, self.one(), self.zero())

    def add_const(self, x, c):
        return self.arithmetic(1, c, x, self.one(), self.one())

    def mul_const_add(self, c, x, y):
        return self.arithmetic(c, 1, x, self.one(), y)

    def not_(self, b):
        return BoolTarget(self.sub(self.one(), b.target))

    def select(self, b, x, y):
        """if b { x } else { y } = b*x - (b*y - y)"""
        tmp = self.mul_sub(b.target, y, y)
        return self.mul_sub(b.target, x, tmp)

    def is_equal(self, x, y):
        zero = self.zero()
        equal = self.add_virtual_bool_target_unsafe()
        inv = self.add_virtual_target()
        self.ops.append((OP_EQ, equal.target, inv, x, y, 0, 0, 0))   # EqualityGenerator
        not_equal = self.not_(equal)
        diff = self.sub(x, y)
        not_equal_check = self.mul(diff, inv)
        diff_normalized = self.mul(diff, equal.target)
        self.connect(not_equal.target, not_equal_check)
        self.connect(diff_normalized, zero)
        return equal

    # ---- Poseidon (gadgets/hash.rs: permute / hash_n_to_hash_no_pad) ----------------------------
    def permute(self, state):
        """One PoseidonGate row (swap wire tied to zero); returns the 12 output wire targets."""
        assert len(state) == 12
        row = self.add_gate(GATE_POSEIDON)
        self.connect(self.zero(), wire(row, 24))
        ins = [wire(row, i) for i in range(12)]
        outs = [wire(row, 12 + i) for i in range(12)]
        for t, w in zip(state, ins):
            self.connect(t, w)
        self.ops.append((OP_POSEIDON, len(self.poseidon_rows), 0, 0, 0, 0, 0, 0))
        self.poseidon_rows.append((row, ins, outs))
        return outs

    def hash_n_to_m_no_pad(self, inputs, m):
        state = [self.zero()] * 12
        for k in range(0, len(inputs), 8):
            chunk = inputs[k:k + 8]
            state = list(chunk) + state[len(chunk):]
            state = self.permute(state)
        out = []
        while True:
            for t in state[:8]:
                out.append(t)
                if len(out) == m:
                    return out
            state = self.permute(state)

    def hash_n_to_hash_no_pad(self, inputs):
        return self.hash_n_to_m_no_pad(inputs, 4)

    # ---- lookups (gadgets/lookup.rs) ---------------------------------------------------------
    def add_lookup_table_from_pairs(self, pairs):
        self.luts.append([(int(a), int(b)) for a, b in pairs])
        self.lut_to_lookups.append([])
        return len(self.luts) - 1

    def add_lookup_from_index(self, looking_in, lut_index):
        out = self.add_virtual_target()
        self.ops.append((OP_LOOKUP, out, looking_in, 0, 0, lut_index, 0, 0))
        self.lut_to_lookups[lut_index].append((looking_in, out))
        return out

    # ---- build ---------------------------------------------------------------------------------
    def _add_all_lookups(self):
        cfg = self.config
        lu_slots = cfg.num_routed_wires // 2
        lut_slots = cfg.num_routed_wires // 3
        lookup_rows, fixed = [], []
        for li, lut in enumerate(self.luts):
            lookups = self.lut_to_lookups[li]
            assert lookups, f"LUT number {li} is unused"
            last_lu_gate = self.num_gates()
            row, slot = None, lu_slots
            for (tin, tout) in lookups:
                if slot == lu_slots:
                    row, slot = self.add_gate(GATE_LOOKUP), 0
                self.connect(wire(row, 2 * slot), tin)
                self.connect(wire(row, 2 * slot + 1), tout)
                slot += 1
            last_lut_gate = self.num_gates()
            # set_lookup_wires pads the last LookupGate with the first LUT entry
            for s in range(slot, lu_slots):
                fixed.append((last_lut_gate - 1, 2 * s, lut[0][0]))
                fixed.append((last_lut_gate - 1, 2 * s + 1, lut[0][1]))
            padding = lu_slots - slot
            num_lut_rows = (len(lut) - 1) // lut_slots + 1
            for _ in range(num_lut_rows):
                self.add_gate(GATE_LOOKUP_TABLE)
            first_lut_gate = self.num_gates() - 1
            self.add_gate(GATE_NOOP)
            mult_pos = []
            for e, (a, b) in enumerate(lut):   # LookupTableGenerator: rows are filled upside down
                r, s = first_lut_gate - e // lut_slots, e % lut_slots
                fixed.append((r, 3 * s, a))
                fixed.append((r, 3 * s + 1, b))
                mult_pos.append((r, 3 * s + 2))
            lookup_rows.append((last_lu_gate, last_lut_gate, first_lut_gate, padding, mult_pos))
        return lookup_rows, fixed

    def build(self, ctx=None, min_degree_bits=0):
        """CircuitBuilder::build::<PoseidonGoldilocksConfig>() -> CircuitData.  `ctx` is the GPU
        context that will hold the preprocessed (constants, sigmas) commitment.  `min_degree_bits` pads with
        NoopGate rows up to 2^min_degree_bits (tests of large degrees without large circuits)."""
        cfg = self.config
        # PublicInputGate row: wires 0..4 are tied to hash_n_to_hash_no_pad(public inputs), computed in
        # circuit by PoseidonGate rows (upstream build() does the same); no inputs -> the zero hash
        pi_hash = self.hash_n_to_hash_no_pad(self.public_inputs) if self.public_inputs else [self.zero()] * 4
        pi_row = self.add_gate(GATE_PUBLIC_INPUT)
        for i in range(4):
            self.connect(wire(pi_row, i), pi_hash[i])
        lookup_rows, fixed = self._add_all_lookups()
        # ConstantGate rows (2 constants each)
        items = list(self.constants_to_targets.items())
        const_ops = []      # constant generators have no inputs: they run first
        for k in range(0, len(items), cfg.num_constants):
            chunk = items[k:k + cfg.num_constants]
            row = self.add_gate(GATE_CONSTANT, [c for c, _ in chunk])
            for i, (c, t) in enumerate(chunk):
                self.connect(t, wire(row, i))
                const_ops.append((OP_CONST, wire(row, i), 0, 0, 0, 0, c, 0))
        self.ops = const_ops + self.ops
        while len(self.rows) < max(4, 1 << min_degree_bits) or (len(self.rows) & (len(self.rows) - 1)):
            self.add_gate(GATE_NOOP)
        n = len(self.rows)
        degree_bits = n.bit_length() - 1
        return CircuitData(self, ctx, degree_bits, lookup_rows, fixed)


def _fri_reduction_arity_bits(degree_bits, fri):
    out, d = [], degree_bits
    while d > fri.final_poly_bits and d + fri.rate_bits - fri.reduction_arity_bits >= fri.cap_height:
        out.append(fri.reduction_arity_bits)
        d -= fri.reduction_arity_bits
    return out


class CircuitData:
    """CircuitData: prover-only + common data.  `prove(pw)` mirrors CircuitData::prove."""

    def __init__(self, b, ctx, degree_bits, lookup_rows, fixed):
        cfg = b.config
        self.config = cfg
        self.ctx = ctx
        self.degree_bits = degree_bits
        n = self.n = 1 << degree_bits
        R = cfg.num_routed_wires
        self.luts = b.luts
        self.lookup_rows = [(a, bb, c) for (a, bb, c, _, _) in lookup_rows]

        # ---- gate set, selectors (gates/selectors.rs) ----
        kinds = sorted({k for k, _ in b.rows}, key=lambda k: GATE_META[k])
        index = {k: i for i, k in enumerate(kinds)}
        max_degree = cfg.max_quotient_degree_factor + 1
        degs = [GATE_META[k][0] for k in kinds]
        if degs[-1] + len(kinds) - 1 <= max_degree:
            groups = [(0, len(kinds))]
        else:
            groups, start = [], 0
            while start < len(kinds):
                size = 0
                while start + size < len(kinds) and size + degs[start + size] < max_degree:
                    size += 1
                groups.append((start, start + size))
                start += size
        sel_of = {}
        for gi, (s, e) in enumerate(groups):
            for i in range(s, e):
                sel_of[i] = gi
        row_gate = np.array([index[k] for k, _ in b.rows], dtype=np.int64)
        selectors = np.zeros((len(groups), n), dtype=np.uint64)
        for gi in range(len(groups)):
            grp = np.array([sel_of[i] for i in row_gate])
            selectors[gi] = np.where(grp == gi, row_gate, UNUSED_SELECTOR if len(groups) > 1 else row_gate).astype(np.uint64)
        self.num_selectors = len(groups)
        # ---- lookup selectors ----
        num_luts = len(b.luts)
        self.num_lookup_selectors = 4 + num_luts if num_luts else 0
        lsel = np.zeros((self.num_lookup_selectors, n), dtype=np.uint64)
        for li, (last_lu, last_lut, first_lut) in enumerate(self.lookup_rows):
            lsel[0, last_lut:first_lut + 1] = 1          # TransSre
            lsel[1, last_lu:last_lut] = 1                # TransLdc
            lsel[2, first_lut + 1] = 1                   # InitSre
            lsel[3, last_lu] = 1                         # LastLdc
            lsel[4 + li, last_lut] = 1                   # StartEnd (one per LUT)
        # ---- gate constants ----
        consts = np.zeros((cfg.num_constants, n), dtype=np.uint64)
        for r, (_, cs) in enumerate(b.rows):
            for i, c in enumerate(cs):
                consts[i, r] = c
        # ---- copy constraints -> partitions, sigma polynomials ----
        nv = b.num_virtual
        total = nv + n * NUM_WIRES          # virtual targets first, then wires

        def tid(t):
            return (-t - 1) if t < 0 else nv + t
        ca = np.array([tid(t) for t in b.copy_a], dtype=np.int64)
        cb = np.array([tid(t) for t in b.copy_b], dtype=np.int64)
        for arr in (np.array(b.copy_a, dtype=np.int64), np.array(b.copy_b, dtype=np.int64)):
            w = arr[arr >= 0]
            assert (w % NUM_WIRES < R).all(), "copy constraint on a non-routed wire"
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        graph = coo_matrix((np.ones(len(ca), dtype=np.int8), (ca, cb)), shape=(total, total))
        self.num_slots, slot_of = connected_components(graph, directed=False)
        self._slot_of = slot_of.astype(np.int32)
        self._nv = nv
        g = pow(1753635133440165772, 1 << (32 - degree_bits), P)
        subgroup = gl_powers(g, n)
        self.k_is = gl_powers(7, R)
        # routed wires grouped by partition; sigma maps each wire to the next one in its cycle
        wrow, wcol = np.meshgrid(np.arange(n), np.arange(R), indexing="ij")
        wt = (wrow * NUM_WIRES + wcol).ravel()            # wire targets, row-major
        wslots = slot_of[nv + wt]
        order = np.argsort(wslots, kind="stable")
        sorted_slots = wslots[order]
        nxt = np.empty_like(order)
        # within each run of equal slot, next = following element, last wraps to first
        starts = np.flatnonzero(np.r_[True, sorted_slots[1:] != sorted_slots[:-1]])
        ends = np.r_[starts[1:], len(order)]
        nxt[order] = np.roll(order, -1)
        nxt[order[ends - 1]] = order[starts]
        tgt_row, tgt_col = wt[nxt] // NUM_WIRES, wt[nxt] % NUM_WIRES
        sig = gl_mul(self.k_is[tgt_col], subgroup[tgt_row])
        sigmas = sig.reshape(n, R).T.copy()               # [R][n]
        self.constants_sigmas = np.ascontiguousarray(np.concatenate([selectors, lsel, consts, sigmas]), dtype=np.uint64)

        # ---- gate table ----
        self.gate_kinds = kinds
        gate_arr = (ffi.Gate * len(kinds))()
        ncons = {GATE_NOOP: 0, GATE_LOOKUP: 0, GATE_LOOKUP_TABLE: 0, GATE_CONSTANT: cfg.num_constants,
                 GATE_PUBLIC_INPUT: 4, GATE_ARITHMETIC: b.ops_per_arith_row, GATE_POSEIDON: 123}
        for i, k in enumerate(kinds):
            s, e = groups[sel_of[i]]
            gate_arr[i] = ffi.Gate(k, sel_of[i], s, e, ncons[k], b.ops_per_arith_row if k == GATE_ARITHMETIC else cfg.num_constants)
        self._gate_arr = gate_arr
        self.num_gate_constraints = max(ncons[k] for k in kinds)
        self.quotient_degree_factor = cfg.max_quotient_degree_factor
        self.num_partial_products = -(-R // self.quotient_degree_factor) - 1
        self.reduction_arity_bits = _fri_reduction_arity_bits(degree_bits, cfg.fri_config)

        # ---- witness program ----
        self.public_input_targets = list(b.public_inputs)
        self._build_witness_program(b, lookup_rows, fixed)
        self.circuit_digest = np.zeros(4, dtype=np.uint64)
        self.constants_sigmas_cap = None
        self._gpu_circuit = None
        self._orc_circuit = None
        if ctx is not None:
            self.load(ctx)

    # -- descriptor shared by the GPU library and (in tests) the oracle: same C layout --
    def descriptor(self):
        cfg = self.config
        d = ffi.CircuitDesc()
        d.degree_bits = self.degree_bits
        d.num_wires, d.num_routed_wires, d.num_constants = cfg.num_wires, cfg.num_routed_wires, cfg.num_constants
        d.num_challenges, d.quotient_degree_factor = cfg.num_challenges, self.quotient_degree_factor
        f = cfg.fri_config
        d.rate_bits, d.cap_height, d.pow_bits, d.num_query_rounds = f.rate_bits, f.cap_height, f.proof_of_work_bits, f.num_query_rounds
        d.num_reduction_arity_bits = len(self.reduction_arity_bits)
        for i, a in enumerate(self.reduction_arity_bits):
            d.reduction_arity_bits[i] = a
        d.num_selectors, d.num_lookup_selectors = self.num_selectors, self.num_lookup_selectors
        d.num_gates, d.gates = len(self.gate_kinds), self._gate_arr
        d.num_gate_constraints = self.num_gate_constraints
        d.num_partial_products = self.num_partial_products
        d.num_luts = len(self.luts)
        self._lut_lens = np.array([len(l) for l in self.luts] or [0], dtype=np.int32)
        self._lut_data = np.array([v for l in self.luts for pr in l for v in pr] or [0], dtype=np.uint16)
        self._lookup_rows_arr = np.array([v for r in self.lookup_rows for v in r] or [0], dtype=np.int32)
        d.lut_lens = self._lut_lens.ctypes.data_as(C.POINTER(C.c_int32))
        d.lut_data = self._lut_data.ctypes.data_as(C.POINTER(C.c_uint16))
        d.lookup_rows = self._lookup_rows_arr.ctypes.data_as(C.POINTER(C.c_int32))
        d.num_public_inputs = len(self.public_input_targets)
        d.k_is = self.k_is.ctypes.data_as(C.POINTER(C.c_uint64))
        d.constants_sigmas = self.constants_sigmas.ctypes.data_as(C.POINTER(C.c_uint64))
        for i in range(4):
            d.circuit_digest[i] = int(self.circuit_digest[i])
        return d

    def load(self, ctx):
        """Upload the preprocessed data: commits (constants, sigmas) on the GPU and derives the
        circuit digest = hash_no_pad(cap || hash_pad([]) || degree_bits) on the device."""
        self.ctx = ctx
        self._wmap = None
        lib = ctx.lib
        f = self.config.fri_config
        from .polynomial_batch import PolynomialBatch
        pre = PolynomialBatch.from_values(ctx, self.constants_sigmas, f.rate_bits, f.cap_height)
        self.constants_sigmas_cap = pre.cap.copy()
        pre.free()
        pad = np.array([[1] + [0] * 10 + [1]], dtype=np.uint64)          # hash_pad(&[])
        dom = ctx.hash_no_pad_many(pad)[0]
        parts = np.concatenate([self.constants_sigmas_cap.ravel(), dom, np.array([self.degree_bits], dtype=np.uint64)])
        self.circuit_digest = ctx.hash_no_pad_many(parts[None, :])[0].copy()
        self._gpu_circuit = self.load_handle(ctx)
        return self

    def load_handle(self, ctx):
        """p2g_circuit_load on another context (a second stream of the same GPU, or another GPU);
        the digest computed by load() is reused."""
        h = C.c_void_p()
        desc = self.descriptor()
        cap = np.empty((1 << self.config.fri_config.cap_height, 4), dtype=np.uint64)
        ctx.check(ctx.lib.p2g_circuit_load(ctx.handle, C.byref(desc), C.byref(h), cap.ctypes.data))
        assert np.array_equal(cap, self.constants_sigmas_cap)
        return h

    @property
    def proof_words(self):
        return self.ctx.lib.p2g_proof_words(self._gpu_circuit)

    # ---- witness ---------------------------------------------------------------------------
    def _slot(self, t):
        s = int(self._slot_of[(-t - 1) if t < 0 else self._nv + t])
        if s >= self.num_active_slots:
            raise ValueError("target is an unconnected wire cell that no generator uses; it cannot be set")
        return s

    def _build_witness_program(self, b, lookup_rows, fixed):
        n = self.n
        so, nv = self._slot_of, self._nv

        def slots(ts):
            ts = np.asarray(ts, dtype=np.int64)
            idx = np.where(ts < 0, -ts - 1, nv + ts)
            return so[idx]
        ops = np.array([o[:6] for o in b.ops], dtype=np.int64).reshape(-1, 6)
        kinds = ops[:, 0]
        prog = np.zeros((len(ops), 6), dtype=np.int32)
        prog[:, 0] = kinds
        prog[:, 1] = slots(ops[:, 1])
        prog[:, 2] = slots(ops[:, 2])
        mp = kinds == OP_POSEIDON
        prog[mp, 1] = ops[mp, 1]            # index into the PoseidonGate row table, not a target
        prog[mp, 2] = 0
        for col in (3, 4):
            m = (kinds == OP_ARITH) | (kinds == OP_EQ)
            prog[m, col] = slots(ops[m, col])
        prog[:, 5] = ops[:, 5]
        # OP_EQ carries (equal, inv, x, y): all four are targets; OP_LOOKUP: (out, in, -, -, lut)
        # OP_CONST: (out) only
        mc = kinds == OP_CONST
        prog[mc, 2] = 0
        # Compact slot numbering: partitions a generator reads or writes, that hold a virtual target or that
        # tie several cells together come first ("active"); the rest are single unconnected wire cells that
        # nothing ever sets (they read as 0, as in PartitionWitness::full_witness).  Only the active part is
        # materialised per witness: 0.31 M instead of 3.7 M values for the AES-GCM circuit of config 2.
        ls_t = [t for l in b.lut_to_lookups for (t, _) in l]
        pos_t = [t for (_, ins, outs) in b.poseidon_rows for t in list(ins) + list(outs)]
        active = np.bincount(so, minlength=self.num_slots) > 1
        active[so[:nv]] = True
        active[prog[~mp, 1]] = True
        active[prog[kinds <= OP_EQ, 2]] = True
        m34 = (kinds == OP_ARITH) | (kinds == OP_EQ)
        active[prog[m34, 3]] = True
        active[prog[m34, 4]] = True
        for ts in (ls_t, pos_t, self.public_input_targets):
            if len(ts):
                active[slots(ts)] = True
        order = np.concatenate([np.flatnonzero(active), np.flatnonzero(~active)])
        renum = np.empty(self.num_slots, dtype=np.int64)
        renum[order] = np.arange(self.num_slots)
        self.num_active_slots = int(active.sum())
        so = self._slot_of = renum[so].astype(np.int32)
        keep = np.ones(len(prog), dtype=bool)
        prog[~mp, 1] = renum[prog[~mp, 1]]
        prog[kinds <= OP_EQ, 2] = renum[prog[kinds <= OP_EQ, 2]]
        prog[m34, 3] = renum[prog[m34, 3]]
        prog[m34, 4] = renum[prog[m34, 4]]
        del keep
        self._w_ops = np.ascontiguousarray(prog)
        self._w_consts = np.array([[o[6], o[7]] for o in b.ops], dtype=np.uint64).reshape(-1, 2)
        # wire -> slot map (column-major); cells whose partition is a singleton never set stay 0
        wt = (np.arange(n)[None, :] * NUM_WIRES + np.arange(NUM_WIRES)[:, None])
        ws = so[nv + wt].astype(np.int32)
        ws[ws >= self.num_active_slots] = -1             # never-set single cells: empty
        self._w_wire_slot = np.ascontiguousarray(ws)
        self._w_fixed_pos = np.array([c * n + r for r, c, _ in fixed] or [0], dtype=np.int64)
        self._w_fixed_val = np.array([v for _, _, v in fixed] or [0], dtype=np.uint64)
        self._w_num_fixed = len(fixed)
        self._w_lookup_counts = np.array([len(l) for l in b.lut_to_lookups] or [0], dtype=np.int32)
        ls = [t for l in b.lut_to_lookups for (t, _) in l]
        self._w_lookup_slots = slots(ls).astype(np.int32) if ls else np.zeros(1, dtype=np.int32)
        self._w_lookup_padding = np.array([r[3] for r in lookup_rows] or [0], dtype=np.int32)
        self._w_mult_pos = np.array([c * n + r for row in lookup_rows for (r, c) in row[4]] or [0], dtype=np.int64)
        pr = [[row] + [int(v) for v in slots(ins)] + [int(v) for v in slots(outs)] for (row, ins, outs) in b.poseidon_rows]
        self._w_poseidon = np.array(pr or [[0] * 25], dtype=np.int32)
        self._w_num_poseidon = len(pr)
        self._w_lut_lens = np.array([len(l) for l in b.luts] or [0], dtype=np.int32)
        self._w_lut_data = np.array([v for l in b.luts for pr in l for v in pr] or [0], dtype=np.uint16)
        self._wprog = None

    def _witness_lib(self):
        # next to the in-tree libp2gpu.so (P2G_LIB_PATH may point at an A/B variant of libp2gpu.so elsewhere)
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libp2witness.so")
        if not os.path.exists(path):
            raise ffi.P2GError(-1, f"{path} not built")
        lib = C.CDLL(path)
        lib.p2w_program_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        lib.p2w_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.p2w_generate_many.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.p2w_set_num_threads.argtypes = [C.c_int32]
        lib.p2w_set_num_threads.restype = None
        lib.p2w_ext_slots.argtypes = [C.c_void_p]
        lib.p2w_ext_slots.restype = C.c_uint32
        lib.p2w_wire_map.argtypes = [C.c_void_p, C.c_void_p]
        lib.p2w_fixed_cells.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        lib.p2w_generate_slots.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.p2w_generate_slots_many.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        return lib

    def _program(self):
        if self._wprog is None:
            class Desc(C.Structure):
                _fields_ = [("num_slots", C.c_uint32), ("num_ops", C.c_uint32), ("ops", C.c_void_p), ("op_consts", C.c_void_p),
                            ("num_luts", C.c_uint32), ("lut_lens", C.c_void_p), ("lut_data", C.c_void_p),
                            ("num_wires", C.c_uint32), ("log_n", C.c_uint32), ("wire_slot", C.c_void_p),
                            ("num_fixed", C.c_uint32), ("fixed_pos", C.c_void_p), ("fixed_val", C.c_void_p),
                            ("lookup_counts", C.c_void_p), ("lookup_slots", C.c_void_p), ("lookup_padding", C.c_void_p),
                            ("mult_pos", C.c_void_p), ("num_poseidon", C.c_uint32), ("poseidon_rows", C.c_void_p)]
            d = Desc(self.num_active_slots, len(self._w_ops), self._w_ops.ctypes.data, self._w_consts.ctypes.data,
                     len(self.luts), self._w_lut_lens.ctypes.data, self._w_lut_data.ctypes.data,
                     NUM_WIRES, self.degree_bits, self._w_wire_slot.ctypes.data,
                     self._w_num_fixed, self._w_fixed_pos.ctypes.data, self._w_fixed_val.ctypes.data,
                     self._w_lookup_counts.ctypes.data, self._w_lookup_slots.ctypes.data, self._w_lookup_padding.ctypes.data,
                     self._w_mult_pos.ctypes.data, self._w_num_poseidon, self._w_poseidon.ctypes.data)
            self._wlib = self._witness_lib()
            self._wdesc = d            # p2w_program_desc (include/p2witness.h); also what p2g_wprog_load takes
            h = C.c_void_p()
            rc = self._wlib.p2w_program_create(C.byref(d), C.byref(h))
            assert rc == 0
            self._wprog = h
        return self._wprog

    def generate_witness(self, pw, out=None):
        """generate_partial_witness + set_lookup_wires + full_witness(): [135][n] wire matrix.
        Raises ValueError when the inputs contradict the circuit (mirrors prove() -> Err,
        /root/reference/aes-gcm/src/circuit_aes.rs:403-405)."""
        prog = self._program()
        slots = np.array([self._slot(t) for t in pw.values], dtype=np.int32)
        vals = np.array(list(pw.values.values()), dtype=np.uint64)
        if out is None:
            out = np.empty((NUM_WIRES, self.n), dtype=np.uint64)
        rc = self._wlib.p2w_generate(prog, slots.ctypes.data, vals.ctypes.data, len(slots), out.ctypes.data)
        if rc != 0:
            raise ValueError({-10: "partition set twice with different values", -11: "lookup input not in table",
                              -12: "generator input unset"}.get(rc, f"witness error {rc}"))
        return out

    # ---- slot form: the device fills the wire matrix (p2g_prove_slots) ---------------------------
    _WERR = {-10: "partition set twice with different values", -11: "lookup input not in table", -12: "generator input unset"}

    @property
    def ext_slots(self):
        """length of the extended slot vector (PartitionWitness::values analogue, include/p2witness.h)"""
        prog = self._program()
        return int(self._wlib.p2w_ext_slots(prog))

    def generate_slots(self, pw, out=None):
        """generate_partial_witness + set_lookup_wires, WITHOUT full_witness(): one value per partition."""
        prog = self._program()
        slots = np.array([self._slot(t) for t in pw.values], dtype=np.int32)
        vals = np.array(list(pw.values.values()), dtype=np.uint64)
        if out is None:
            out = np.empty(self.ext_slots, dtype=np.uint64)
        rc = self._wlib.p2w_generate_slots(prog, slots.ctypes.data, vals.ctypes.data, len(slots), out.ctypes.data)
        if rc != 0:
            raise ValueError(self._WERR.get(rc, f"witness error {rc}"))
        return out

    def generate_slots_many(self, targets, values, out=None):
        """Batch form of generate_slots: [count][ext_slots]."""
        prog = self._program()
        slots = np.array([self._slot(t) for t in targets], dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.uint64)
        count = values.shape[0]
        if out is None:
            out = np.empty((count, self.ext_slots), dtype=np.uint64)
        rc = self._wlib.p2w_generate_slots_many(prog, slots.ctypes.data, values.ctypes.data, len(slots), count, out.ctypes.data)
        if rc != 0:
            raise ValueError(self._WERR.get(rc, f"witness error {rc}"))
        return out

    def public_inputs_of(self, slots):
        """values of the registered public inputs in a slot vector (PartitionWitness::get_targets)"""
        return np.array([slots[self._slot(t)] for t in self.public_input_targets], dtype=np.uint64)

    def load_wire_map(self, ctx=None, circuit=None):
        """Uploads the wire map (representative_map analogue) next to a loaded circuit; returns the handle."""
        ctx = ctx or self.ctx
        circuit = circuit or self._gpu_circuit
        if ctx is None or circuit is None:
            raise ffi.P2GError(-1, "circuit not loaded on a GPU context (no CPU fallback)")
        prog = self._program()
        wm = np.empty((NUM_WIRES, self.n), dtype=np.int32)
        assert self._wlib.p2w_wire_map(prog, wm.ctypes.data) == 0
        cnt, pos, val = C.c_uint32(), C.c_void_p(), C.c_void_p()
        assert self._wlib.p2w_fixed_cells(prog, C.byref(cnt), C.byref(pos), C.byref(val)) == 0
        h = C.c_void_p()
        ctx.check(ctx.lib.p2g_wmap_load(ctx.handle, circuit, wm.ctypes.data, self.ext_slots, pos, val, cnt.value, C.byref(h)))
        return h

    # ---- input form: the device runs the generators too (p2g_wprog_load / p2g_prove_inputs) ---------
    def load_witness_program(self, ctx, targets):
        """Uploads the level-scheduled generator program for witnesses given by the values of `targets`
        (the PartialWitness::set_target calls of the caller, fixed per program); returns the handle."""
        self._program()
        slots = np.array([self._slot(t) for t in targets], dtype=np.int32)
        h = C.c_void_p()
        ctx.check(ctx.lib.p2g_wprog_load(ctx.handle, C.byref(self._wdesc), slots.ctypes.data, len(slots), C.byref(h)))
        assert ctx.lib.p2g_wprog_ext_slots(h) == self.ext_slots
        return h

    def generate_slots_device(self, ctx, wprog, values):
        """device twin of generate_slots_many: [count][len(targets)] input values -> [count][ext_slots]"""
        values = np.ascontiguousarray(values, dtype=np.uint64)
        out = np.empty((values.shape[0], self.ext_slots), dtype=np.uint64)
        rc = ctx.lib.p2g_wprog_generate(ctx.handle, wprog, values.ctypes.data, values.shape[0], out.ctypes.data)
        if rc in self._WERR:
            raise ValueError(self._WERR[rc])
        ctx.check(rc)
        return out

    def prove_inputs(self, values, wprog, ctx=None, circuit=None, wmap=None, public_inputs=None):
        """input values -> proof: generators, full_witness and the prover all on the device"""
        ctx = ctx or self.ctx
        circuit = circuit or self._gpu_circuit
        if wmap is None:
            if getattr(self, "_wmap", None) is None:
                self._wmap = self.load_wire_map(ctx, circuit)
            wmap = self._wmap
        words = self.proof_words
        out = np.empty(words, dtype=np.uint64)
        got = C.c_size_t()
        values = np.ascontiguousarray(values, dtype=np.uint64)
        pi = np.ascontiguousarray(public_inputs, dtype=np.uint64) if public_inputs is not None and len(public_inputs) else None
        rc = ctx.lib.p2g_prove_inputs(ctx.handle, circuit, wmap, wprog, values.ctypes.data, pi.ctypes.data if pi is not None else None,
                                      out.ctypes.data, words, C.byref(got))
        if rc in self._WERR:
            raise ValueError(self._WERR[rc])
        ctx.check(rc)
        return out[:got.value]

    def prove_slots(self, slots, ctx=None, circuit=None, wmap=None, public_inputs=None):
        """Slot vector -> proof; the wire matrix is gathered on the device (PartitionWitness::full_witness)."""
        ctx = ctx or self.ctx
        circuit = circuit or self._gpu_circuit
        if ctx is None or circuit is None:
            raise ffi.P2GError(-1, "circuit not loaded on a GPU context (no CPU fallback)")
        if wmap is None:
            if getattr(self, "_wmap", None) is None:
                self._wmap = self.load_wire_map(ctx, circuit)
            wmap = self._wmap
        words = self.proof_words
        out = np.empty(words, dtype=np.uint64)
        got = C.c_size_t()
        slots = np.ascontiguousarray(slots, dtype=np.uint64)
        if public_inputs is None and self.public_input_targets:
            public_inputs = self.public_inputs_of(slots)
        pi = np.ascontiguousarray(public_inputs, dtype=np.uint64) if public_inputs is not None and len(public_inputs) else None
        rc = ctx.lib.p2g_prove_slots(ctx.handle, circuit, wmap, slots.ctypes.data, pi.ctypes.data if pi is not None else None,
                                     out.ctypes.data, words, C.byref(got))
        if rc == -3:
            raise ValueError("witness does not satisfy the circuit (P2G_E_UNSAT)")
        ctx.check(rc)
        return out[:got.value]

    def generate_witnesses(self, targets, values, out=None):
        """Batch form: `targets` (list) and `values` [count][len(targets)] -> [count][135][n]."""
        prog = self._program()
        slots = np.array([self._slot(t) for t in targets], dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.uint64)
        count = values.shape[0]
        if out is None:
            out = np.empty((count, NUM_WIRES, self.n), dtype=np.uint64)
        rc = self._wlib.p2w_generate_many(prog, slots.ctypes.data, values.ctypes.data, len(slots), count, out.ctypes.data)
        if rc != 0:
            raise ValueError(f"witness error {rc}")
        return out

    # ---- prove ------------------------------------------------------------------------------
    def prove(self, pw):
        """CircuitData::prove(pw) -> proof (flat u64 words, layout in DESIGN.md)."""
        return self.prove_slots(self.generate_slots(pw))

    def prove_wires(self, wires, public_inputs=None):
        ctx = self.ctx
        if ctx is None or self._gpu_circuit is None:
            raise ffi.P2GError(-1, "circuit not loaded on a GPU context (no CPU fallback)")
        words = self.proof_words
        out = np.empty(words, dtype=np.uint64)
        got = C.c_size_t()
        wires = np.ascontiguousarray(wires, dtype=np.uint64)
        assert (public_inputs is not None and len(public_inputs) == len(self.public_input_targets)) or not self.public_input_targets
        pi = np.ascontiguousarray(public_inputs, dtype=np.uint64) if self.public_input_targets else None
        rc = ctx.lib.p2g_prove(ctx.handle, self._gpu_circuit, wires.ctypes.data, pi.ctypes.data if pi is not None else None,
                               out.ctypes.data, words, C.byref(got))
        if rc == -3:
            raise ValueError("witness does not satisfy the circuit (P2G_E_UNSAT)")
        ctx.check(rc)
        return out[:got.value]
