#!/usr/bin/env python3
"""Emit csrc/das_mimo_vm.inc: the inline-PTX microphone loop of das_mimo_kernel.

The loop over one chunk of (group, microphone) entries is a threaded interpreter: every entry
carries the index of its handler (uniform delay / delay changes once at direction s = 1..7 /
general / end of chunk) and every handler ends with `brx.idx.uni` straight into the handler of
the NEXT entry, whose word was loaded while the current add burst ran.  One indirect branch per
entry, no divergence bookkeeping (BSSY/BSYNC), and the entry decode overlaps the adds because
each handler is one basic block that ptxas schedules freely.

Shipped schedules (measured on B200, C3, 128 frames per launch; profiles/r2_kernel_variants.md): pad uses
carry + fastbr (13 244 maps/s against 11 978 without either), lerp fastbr only (6 127 against 5 713 with
carry); `weave` and `late` measured no better and are kept as generator options only.

Arithmetic is unchanged: per direction the microphones are added in table order with
add.rn.f32x2 (pad, pad_and_sum.c:45) or fma.rn.f32x2 + add.rn.f32x2 (lerp, lerp_and_sum.c:54).
SHARED variants (tolerance mode, exact_sum = 2) add uniform entries once per group into a shared
accumulator row instead of into each of the 8 directions (different rounding order).

    python tools/gen_mimo_vm.py        # rewrites the .inc next to das_mimo.cu
"""
import os

R = 8
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "zybo-rt-sampler-image-detection_b200", "csrc", "das_mimo_vm.inc")


def gen(J, lerp, shared, carry=False, fastbr=False, weave=False, late=False, pair=False):
    """PTX text of one chunk loop.  Operands: %0 .. %(R*Q-1) accumulators (r*Q+q, .b64),
    then [U_0..U_{Q-1} if shared], then eptr (+r), rowp (+r), micb (r), rowb (r).

    carry:  the a-rows of entry i+1 are loaded at the END of handler i (into the registers the adds of
            entry i just released), so that their latency overlaps the indirect branch instead of
            following it; every handler then starts with its first operand rows already in flight.
    fastbr: a predicated direct branch takes the (78 %) uniform case; only the rest pays the indexed one.
    weave:  (with carry) the next entry's row pair q is loaded right behind the adds that consumed pair q of
            the current entry, in the order the next handler will use them, instead of all at the end.
    late:   (without carry) row pair q is loaded just before its adds instead of all rows up front --
            fewer live registers (lerp at 96 registers).
    pair:   (pad; measured, NOT shipped) handler code 10 = this entry AND the next one are uniform: both rows are
            added in one handler (same order per accumulator), one decode and one branch for two entries; needs a
            table builder that marks the pairs greedily inside every chunk.  12 333 against 12 796 maps/s without
            the marks on the same box: ptxas gives the 64-add block its own accumulator allocation and ends it
            with 28 register moves, and wraps the second predicated branch of every dispatch in BSSY / BSYNC."""
    assert not (pair and lerp)
    Q = J // 2
    nacc = R * Q
    nU = Q if shared else 0
    o_eptr, o_rowp, o_micb, o_rowb = (f"%{nacc + nU + i}" for i in range(4))
    acc = lambda r, q: f"%{r * Q + q}"
    U = lambda q: f"%{nacc + q}"
    estride = 48 if lerp else 16
    L = []
    a = L.append
    a("{")
    a(".reg .b32 ex, ey, hc, pa, pb, gx, gy, gz, gw, pg;")
    a(".reg .pred pu;")
    a(f".reg .b32 ta<{J}>, tb<{J}>;")
    a(f".reg .b64 A<{Q}>, B<{Q}>;")
    if lerp:
        a(f".reg .b32 da<{J}>, db<{J}>;")
        a(f".reg .b64 DA<{Q}>, DB<{Q}>, T;")
        a(f".reg .b32 w<{R}>;")
        a(f".reg .b64 W<{R}>;")
    if pair:
        a(".reg .pred pv;")
    a("TS: .branchtargets HU, H1, H2, H3, H4, H5, H6, H7, HG, HEND%s;" % (", HV" if pair else ""))

    def load_rows(ptr, t, d):
        for j in range(J):
            a(f"ld.shared.f32 {t}{j}, [{ptr}+{(j // 2) * 256 + (j % 2) * 128}];")
        if lerp:
            a(f"add.u32 pg, {ptr}, {o_rowb};")
            for j in range(J):
                a(f"ld.shared.f32 {d}{j}, [pg+{(j // 2) * 256 + (j % 2) * 128}];")

    def pack(t, P, d=None, DP=None):
        for q in range(Q):
            a(f"mov.b64 {P}{q}, {{{t}{2 * q}, {t}{2 * q + 1}}};")
            if lerp:
                a(f"mov.b64 {DP}{q}, {{{d}{2 * q}, {d}{2 * q + 1}}};")

    def load_weights():
        if not lerp:
            return
        a(f"ld.shared.v4.f32 {{w0, w1, w2, w3}}, [{o_eptr}+16];")
        a(f"ld.shared.v4.f32 {{w4, w5, w6, w7}}, [{o_eptr}+32];")

    def pack_weights():
        if lerp:
            for r in range(R):
                a(f"mov.b64 W{r}, {{w{r}, w{r}}};")

    def advance():
        a(f"add.u32 {o_eptr}, {o_eptr}, {estride};")
        a(f"add.u32 {o_rowp}, {o_rowp}, {o_micb};")
        a(f"ld.shared.v2.u32 {{ex, ey}}, [{o_eptr}];")

    def accumulate(r, q, P, DP):
        if lerp:
            a(f"fma.rn.f32x2 T, W{r}, {DP}{q}, {P}{q};")
            a(f"add.rn.f32x2 {acc(r, q)}, {acc(r, q)}, T;")
        else:
            a(f"add.rn.f32x2 {acc(r, q)}, {acc(r, q)}, {P}{q};")

    def load_pair(ptr, t, d, q):
        for j in (2 * q, 2 * q + 1):
            a(f"ld.shared.f32 {t}{j}, [{ptr}+{(j // 2) * 256 + (j % 2) * 128}];")
        if lerp:
            for j in (2 * q, 2 * q + 1):
                a(f"ld.shared.f32 {d}{j}, [pg+{(j // 2) * 256 + (j % 2) * 128}];")

    def pack_pair(t, P, d, DP, q):
        a(f"mov.b64 {P}{q}, {{{t}{2 * q}, {t}{2 * q + 1}}};")
        if lerp:
            a(f"mov.b64 {DP}{q}, {{{d}{2 * q}, {d}{2 * q + 1}}};")

    def first_rows():
        """a-rows (first offset of the entry word in ex) of the entry at rowp -> ta / da"""
        a("and.b32 pa, ex, 0xffff;")
        a(f"add.u32 pa, pa, {o_rowp};")
        load_rows("pa", "ta", "da")

    def next_ptr():
        """pointer to the a-rows of the entry whose word is in ex (after advance())"""
        a("and.b32 pa, ex, 0xffff;")
        a(f"add.u32 pa, pa, {o_rowp};")
        if lerp:
            a(f"add.u32 pg, pa, {o_rowb};")

    def dispatch(rows_done=False):
        if carry and not rows_done:
            first_rows()
        a("and.b32 hc, ey, 0xff;")
        if fastbr:
            if pair:
                a("setp.eq.u32 pv, hc, 10;")
                a("@pv bra.uni HV;")
            a("setp.eq.u32 pu, hc, 0;")
            a("@pu bra.uni HU;")
        a("brx.idx.uni hc, TS;")

    a(f"ld.shared.v2.u32 {{ex, ey}}, [{o_eptr}];")
    dispatch()

    # ---- uniform: one delay for all 8 directions ----------------------------------------
    def uniform_adds(q, P="A", DP="DA"):
        if shared:
            a(f"add.rn.f32x2 {U(q)}, {U(q)}, {P}{q};")
            if lerp:
                for r in range(R):
                    a(f"fma.rn.f32x2 {acc(r, q)}, W{r}, {DP}{q}, {acc(r, q)};")
        else:
            for r in range(R):
                accumulate(r, q, P, DP)

    a("HU:")
    if carry and weave:
        load_weights()
        pack("ta", "A", "da", "DA")
        pack_weights()
        advance()
        lead = min(2, Q)                  # the next entry's word needs ~2 pairs of adds to arrive
        for q in range(lead):
            uniform_adds(q)
        next_ptr()
        for q in range(lead):
            load_pair("pa", "ta", "da", q)
        for q in range(lead, Q):
            uniform_adds(q)
            load_pair("pa", "ta", "da", q)
        dispatch(rows_done=True)
    elif late and not carry:
        a("and.b32 pa, ex, 0xffff;")
        a(f"add.u32 pa, pa, {o_rowp};")
        if lerp:
            a(f"add.u32 pg, pa, {o_rowb};")
        load_weights()
        pack_weights()
        advance()
        for q in range(Q):
            load_pair("pa", "ta", "da", q)
            pack_pair("ta", "A", "da", "DA", q)
            uniform_adds(q)
        dispatch()
    else:
        if not carry:
            first_rows()
        load_weights()
        pack("ta", "A", "da", "DA")
        pack_weights()
        advance()
        for q in range(Q):
            uniform_adds(q)
        dispatch()

    # ---- two uniform entries in a row (pad): rows of entry i in ta, of entry i + 1 in tb --------
    if pair:
        a("HV:")
        if not carry:
            first_rows()
        pack("ta", "A")
        advance()                                   # word of entry i + 1 (uniform: only its offset is used)
        a("and.b32 pb, ex, 0xffff;")
        a(f"add.u32 pb, pb, {o_rowp};")
        load_rows("pb", "tb", "db")
        for q in range(Q):
            uniform_adds(q)
        pack("tb", "B")
        advance()                                   # word of entry i + 2
        for q in range(Q):
            uniform_adds(q, "B", "DB")
        dispatch()

    # ---- two runs: directions [0, s) use delay a, [s, 8) delay b ------------------------
    for s in range(1, R):
        a(f"H{s}:")
        if late and not carry:
            a("and.b32 pa, ex, 0xffff;")
            a(f"add.u32 pa, pa, {o_rowp};")
            a("shr.u32 pb, ex, 16;")
            a(f"add.u32 pb, pb, {o_rowp};")
            load_weights()
            pack_weights()
            advance()
            for (ptr, t, d, P, DP, rr) in (("pa", "ta", "da", "A", "DA", range(s)), ("pb", "tb", "db", "B", "DB", range(s, R))):
                if lerp:
                    a(f"add.u32 pg, {ptr}, {o_rowb};")
                for q in range(Q):
                    load_pair(ptr, t, d, q)
                    pack_pair(t, P, d, DP, q)
                    for r in rr:
                        accumulate(r, q, P, DP)
            dispatch()
            continue
        if not carry:
            first_rows()
        a("shr.u32 pb, ex, 16;")
        a(f"add.u32 pb, pb, {o_rowp};")
        load_rows("pb", "tb", "db")
        load_weights()
        pack("ta", "A", "da", "DA")
        pack("tb", "B", "db", "DB")
        pack_weights()
        advance()
        for q in range(Q):
            for r in range(s):
                accumulate(r, q, "A", "DA")
        for q in range(Q):
            for r in range(s, R):
                accumulate(r, q, "B", "DB")
        dispatch()

    # ---- general: every direction its own delay (0.4 % of the C3 table) ------------------
    a("HG:")
    a(f"ld.shared.v4.u32 {{gx, gy, gz, gw}}, [{o_eptr}];")
    load_weights()
    pack_weights()
    offs = ["and.b32 pa, gx, 0xffff;", "shr.u32 pa, gx, 16;",
            "and.b32 pa, gz, 0xffff;", "shr.u32 pa, gz, 16;",
            "and.b32 pa, gw, 0xffff;", "shr.u32 pa, gw, 16;",
            "bfe.u32 pa, gy, 8, 12;", "shr.u32 pa, gy, 20;"]
    for r in range(R):
        if not (carry and r == 0):
            a(offs[r])
            a(f"add.u32 pa, pa, {o_rowp};")
            load_rows("pa", "ta", "da")
        pack("ta", "A", "da", "DA")
        for q in range(Q):
            accumulate(r, q, "A", "DA")
    advance()
    dispatch()
    a("HEND:")
    a("}")
    return L


def c_string(lines):
    return "\n".join('    "%s\\n"' % ln for ln in lines)


def main():
    out = ["// das_mimo_vm.inc -- GENERATED by tools/gen_mimo_vm.py; do not edit.",
           "// Inline-PTX chunk loop (threaded interpreter) of das_mimo_kernel; see the generator's docstring.",
           ""]
    for J in (2, 4, 8):
        Q = J // 2
        for lerp in (0, 1):
            for shared in (0, 1):
                name = f"BF_VM_{'LERP' if lerp else 'PAD'}{'_SHARED' if shared else ''}_J{J}"
                out.append(f"#define {name} \\")
                body = c_string(gen(J, bool(lerp), bool(shared), carry=not lerp, fastbr=True))
                out.append(" \\\n".join(body.split("\n")))
                out.append("")
        # operand list of the accumulators: acc is unsigned long long [8][Q]
        ops = ", ".join(f'"+l"(ACC[{r}][{q}])' for r in range(R) for q in range(Q))
        out.append(f"#define BF_VM_ACC_OPERANDS_J{J}(ACC) {ops}")
        uops = ", ".join(f'"+l"(U[{q}])' for q in range(Q))
        out.append(f"#define BF_VM_U_OPERANDS_J{J}(U) {uops}")
        out.append("")
    with open(OUT, "w") as f:
        f.write("\n".join(out))
    print(OUT)


if __name__ == "__main__":
    main()
