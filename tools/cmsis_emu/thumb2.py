"""A small ARMv7E-M (Thumb-2 + VFPv4-SP) interpreter, enough to run leaf DSP routines of a static archive.

Purpose (test infrastructure, container only): the reference vendors CMSIS-DSP V1.4.5b only as an ARM archive
(receiver/Drivers/CMSIS/Lib/libarm_cortexM4lf_math.a).  No ARM toolchain or emulator exists in this image, so this
interpreter links the archive's members itself and executes the reference's OWN machine code for arm_rfft_fast_f32,
arm_cfft_f32, arm_cmplx_mult_cmplx_f32, arm_fir_f32, ... on host-provided buffers.  tools/cmsis_emu/make_vectors.py
uses it to write tests/golden/cmsis_binary_vectors.npz: input/output pairs of the real reference arithmetic that pin
the oracle where no device capture does.

Scope: user-mode integer Thumb/Thumb-2 subset that GCC 5.4 -O3 emits for these routines, single-precision VFP with
round-to-nearest (the Cortex-M4 reset state: FPSCR.RMode = RN, FZ = 0, DN = 0).  Unknown encodings raise.
"""
import struct

import numpy as np

M32 = 0xFFFFFFFF


class EmuError(RuntimeError):
    pass


def _sx(v, bits):
    return v - (1 << bits) if v & (1 << (bits - 1)) else v


def _f(u):
    return np.frombuffer(struct.pack("<I", u & M32), dtype=np.float32)[0]


def _u(f):
    return struct.unpack("<I", np.float32(f).tobytes())[0]


def thumb_expand_imm(imm12, carry):
    """ThumbExpandImm_C -> (value, carry)"""
    if (imm12 >> 10) == 0:
        b = imm12 & 0xFF
        m = (imm12 >> 8) & 3
        if m == 0:
            v = b
        elif m == 1:
            v = (b << 16) | b
        elif m == 2:
            v = (b << 24) | (b << 8)
        else:
            v = (b << 24) | (b << 16) | (b << 8) | b
        return v, carry
    un = 0x80 | (imm12 & 0x7F)
    rot = imm12 >> 7
    v = ((un >> rot) | (un << (32 - rot))) & M32
    return v, v >> 31


class Cpu:
    RETURN_MAGIC = 0xFFFFFF00

    def __init__(self, mem_size=1 << 24):
        self.mem = bytearray(mem_size)
        self.r = [0] * 16
        self.s = [0] * 32                       # raw bits of s0..s31
        self.n = self.z = self.c = self.v = 0
        self.fn = self.fz = self.fc = self.fv = 0
        self.it = []                            # remaining conditions of the current IT block
        self.icount = 0
        self.hooks = {}                         # address -> python callable standing in for a libc routine
        np.seterr(all="ignore")

    # ---- memory ----
    def rd32(self, a):
        return struct.unpack_from("<I", self.mem, a)[0]

    def rd16(self, a):
        return struct.unpack_from("<H", self.mem, a)[0]

    def rd8(self, a):
        return self.mem[a]

    def wr32(self, a, v):
        struct.pack_into("<I", self.mem, a, v & M32)

    def wr16(self, a, v):
        struct.pack_into("<H", self.mem, a, v & 0xFFFF)

    def wr8(self, a, v):
        self.mem[a] = v & 0xFF

    # ---- flags / ALU helpers ----
    def cond(self, c):
        n, z, cc, v = self.n, self.z, self.c, self.v
        base = (z, cc, n, v, cc and not z, n == v, (n == v) and not z, 1)[c >> 1]
        base = 1 if base else 0
        if (c & 1) and c != 15:
            base ^= 1
        return base

    def add_c(self, a, b, cin, setf):
        full = a + b + cin
        res = full & M32
        if setf:
            self.n = res >> 31
            self.z = int(res == 0)
            self.c = int(full > M32)
            self.v = int((_sx(a, 32) + _sx(b, 32) + cin) != _sx(res, 32))
        return res

    def nz(self, res, carry=None):
        self.n = res >> 31
        self.z = int(res == 0)
        if carry is not None:
            self.c = carry

    def shift_c(self, val, typ, amt, cin):
        """typ 0 LSL, 1 LSR, 2 ASR, 3 ROR (amt already decoded; RRX not supported) -> (res, carry)"""
        if amt == 0:
            return val, cin
        if typ == 0:
            if amt > 32:
                return 0, 0
            full = val << amt
            return full & M32, (full >> 32) & 1
        if typ == 1:
            if amt > 32:
                return 0, 0
            return (val >> amt) & M32, (val >> (amt - 1)) & 1
        if typ == 2:
            sv = _sx(val, 32)
            amt = min(amt, 32)
            return (sv >> amt) & M32, (sv >> (amt - 1)) & 1
        amt &= 31
        if amt == 0:
            return val, val >> 31
        res = ((val >> amt) | (val << (32 - amt))) & M32
        return res, res >> 31

    @staticmethod
    def decode_imm_shift(typ, imm5):
        if typ in (1, 2) and imm5 == 0:
            return typ, 32
        if typ == 3 and imm5 == 0:
            raise EmuError("RRX")
        return typ, imm5

    def pc_read(self):
        return (self.r[15] + 4) & M32          # r[15] holds the address of the current instruction

    def reg(self, i):
        return self.pc_read() if i == 15 else self.r[i]

    def branch(self, target):
        self.next_pc = target & ~1 & M32

    def dp_op(self, op, rn_val, op2, carry, setf, rd):
        """data-processing opcodes of the Thumb-2 32-bit encodings; returns result or None (compare)"""
        if op == 0:      # AND / TST
            res = rn_val & op2
            if setf:
                self.nz(res, carry)
            return None if rd == 15 and setf else res
        if op == 1:      # BIC
            res = rn_val & ~op2 & M32
        elif op == 2:    # ORR / MOV
            res = op2 if rn_val is None else (rn_val | op2)
        elif op == 3:    # ORN / MVN
            res = (~op2 & M32) if rn_val is None else (rn_val | (~op2 & M32))
        elif op == 4:    # EOR / TEQ
            res = rn_val ^ op2
            if setf:
                self.nz(res, carry)
            return None if rd == 15 and setf else res
        elif op == 8:    # ADD / CMN
            res = self.add_c(rn_val, op2, 0, setf)
            return None if rd == 15 and setf else res
        elif op == 10:   # ADC
            return self.add_c(rn_val, op2, self.c, setf)
        elif op == 11:   # SBC
            return self.add_c(rn_val, ~op2 & M32, self.c, setf)
        elif op == 13:   # SUB / CMP
            res = self.add_c(rn_val, ~op2 & M32, 1, setf)
            return None if rd == 15 and setf else res
        elif op == 14:   # RSB
            return self.add_c(~rn_val & M32, op2, 1, setf)
        else:
            raise EmuError("dp op %d" % op)
        if setf:
            self.nz(res, carry)
        return res

    # ---- execution ----
    def call(self, addr, args=(), max_instr=200_000_000):
        for i, a in enumerate(args):
            self.r[i] = a & M32
        self.r[14] = self.RETURN_MAGIC | 1
        self.r[15] = addr & ~1
        self.it = []
        start = self.icount
        while self.r[15] != self.RETURN_MAGIC:
            self.step()
            if self.icount - start > max_instr:
                raise EmuError("instruction budget exceeded")
        return self.r[0]

    def step(self):
        pc = self.r[15]
        if pc in self.hooks:                    # a host routine: run it and return to the caller
            self.hooks[pc](self)
            self.icount += 1
            self.branch_x(self.r[14])
            self.r[15] = self.next_pc
            return
        hw = self.rd16(pc)
        wide = (hw >> 11) >= 0x1D
        self.next_pc = pc + (4 if wide else 2)
        self.icount += 1
        execute = True
        in_it = bool(self.it)
        if in_it:
            execute = bool(self.cond(self.it.pop(0)))
        if execute:
            if wide:
                self.exec32(hw, self.rd16(pc + 2), in_it)
            else:
                self.exec16(hw, in_it)
        self.r[15] = self.next_pc

    # ---- 16-bit ----
    def exec16(self, hw, in_it):
        r = self.r
        setf = not in_it
        top = hw >> 10
        if top < 0x10:
            op = hw >> 11
            if op < 3:                                            # LSL/LSR/ASR imm
                imm5, rm, rd = (hw >> 6) & 31, (hw >> 3) & 7, hw & 7
                if op == 0 and imm5 == 0:                         # MOVS rd, rm
                    res, cy = r[rm], self.c
                else:
                    typ, amt = self.decode_imm_shift(op, imm5)
                    res, cy = self.shift_c(r[rm], typ, amt, self.c)
                r[rd] = res
                if setf:
                    self.nz(res, cy)
                return
            if op == 3:
                sub = (hw >> 9) & 3
                rn, rd = (hw >> 3) & 7, hw & 7
                val = r[(hw >> 6) & 7] if sub < 2 else (hw >> 6) & 7
                r[rd] = self.add_c(r[rn], val, 0, setf) if (sub & 1) == 0 else self.add_c(r[rn], ~val & M32, 1, setf)
                return
            rd, imm8 = (hw >> 8) & 7, hw & 0xFF
            if op == 4:
                r[rd] = imm8
                if setf:
                    self.nz(imm8)
            elif op == 5:
                self.add_c(r[rd], ~imm8 & M32, 1, True)
            elif op == 6:
                r[rd] = self.add_c(r[rd], imm8, 0, setf)
            else:
                r[rd] = self.add_c(r[rd], ~imm8 & M32, 1, setf)
            return
        if top == 0x10:                                           # data processing
            op, rm, rdn = (hw >> 6) & 15, (hw >> 3) & 7, hw & 7
            a, b = r[rdn], r[rm]
            if op == 0:
                res = a & b
            elif op == 1:
                res = a ^ b
            elif op in (2, 3, 4, 7):
                typ = {2: 0, 3: 1, 4: 2, 7: 3}[op]
                res, cy = self.shift_c(a, typ, b & 0xFF, self.c)
                r[rdn] = res
                if setf:
                    self.nz(res, cy)
                return
            elif op == 5:
                r[rdn] = self.add_c(a, b, self.c, setf)
                return
            elif op == 6:
                r[rdn] = self.add_c(a, ~b & M32, self.c, setf)
                return
            elif op == 8:
                self.nz(a & b)
                return
            elif op == 9:
                r[rdn] = self.add_c(~b & M32, 0, 1, setf)
                return
            elif op == 10:
                self.add_c(a, ~b & M32, 1, True)
                return
            elif op == 11:
                self.add_c(a, b, 0, True)
                return
            elif op == 12:
                res = a | b
            elif op == 13:
                res = (a * b) & M32
            elif op == 14:
                res = a & ~b & M32
            else:
                res = ~b & M32
            r[rdn] = res
            if setf:
                self.nz(res)
            return
        if top == 0x11:                                           # special data / branch exchange
            op = (hw >> 8) & 3
            rm = (hw >> 3) & 15
            rd = (hw & 7) | ((hw >> 4) & 8)
            if op == 0:
                res = (self.reg(rd) + self.reg(rm)) & M32
                if rd == 15:
                    self.branch(res)
                else:
                    r[rd] = res
            elif op == 1:
                self.add_c(self.reg(rd), ~self.reg(rm) & M32, 1, True)
            elif op == 2:
                if rd == 15:
                    self.branch(self.reg(rm))
                else:
                    r[rd] = self.reg(rm)
            else:
                tgt = self.reg(rm)
                if hw & 0x80:
                    r[14] = (self.next_pc | 1) & M32
                self.branch_x(tgt)
            return
        if (hw >> 11) == 9:                                       # LDR literal
            rt, imm = (hw >> 8) & 7, (hw & 0xFF) * 4
            r[rt] = self.rd32((self.pc_read() & ~3) + imm)
            return
        if (hw >> 12) == 5:                                       # load/store register offset
            op, rm, rn, rt = (hw >> 9) & 7, (hw >> 6) & 7, (hw >> 3) & 7, hw & 7
            a = (r[rn] + r[rm]) & M32
            if op == 0:
                self.wr32(a, r[rt])
            elif op == 1:
                self.wr16(a, r[rt])
            elif op == 2:
                self.wr8(a, r[rt])
            elif op == 3:
                r[rt] = _sx(self.rd8(a), 8) & M32
            elif op == 4:
                r[rt] = self.rd32(a)
            elif op == 5:
                r[rt] = self.rd16(a)
            elif op == 6:
                r[rt] = self.rd8(a)
            else:
                r[rt] = _sx(self.rd16(a), 16) & M32
            return
        if (hw >> 13) == 3:                                       # STR/LDR/STRB/LDRB imm5
            byte, load = (hw >> 12) & 1, (hw >> 11) & 1
            imm5, rn, rt = (hw >> 6) & 31, (hw >> 3) & 7, hw & 7
            a = (r[rn] + (imm5 if byte else imm5 * 4)) & M32
            if load:
                r[rt] = self.rd8(a) if byte else self.rd32(a)
            elif byte:
                self.wr8(a, r[rt])
            else:
                self.wr32(a, r[rt])
            return
        if (hw >> 12) == 8:                                       # STRH/LDRH imm5
            load, imm5, rn, rt = (hw >> 11) & 1, (hw >> 6) & 31, (hw >> 3) & 7, hw & 7
            a = (r[rn] + imm5 * 2) & M32
            if load:
                r[rt] = self.rd16(a)
            else:
                self.wr16(a, r[rt])
            return
        if (hw >> 12) == 9:                                       # STR/LDR sp-relative
            load, rt, imm = (hw >> 11) & 1, (hw >> 8) & 7, (hw & 0xFF) * 4
            a = (r[13] + imm) & M32
            if load:
                r[rt] = self.rd32(a)
            else:
                self.wr32(a, r[rt])
            return
        if (hw >> 12) == 10:                                      # ADR / ADD rd, sp, imm
            rd, imm = (hw >> 8) & 7, (hw & 0xFF) * 4
            r[rd] = ((r[13] if hw & 0x800 else (self.pc_read() & ~3)) + imm) & M32
            return
        if (hw >> 12) == 11:                                      # miscellaneous
            sub = (hw >> 8) & 15
            if sub == 0:
                imm = (hw & 0x7F) * 4
                r[13] = (r[13] - imm if hw & 0x80 else r[13] + imm) & M32
            elif sub in (1, 3, 9, 11):                            # CBZ / CBNZ
                rn = hw & 7
                imm = ((hw >> 3) & 0x1F) * 2 + ((hw >> 9) & 1) * 64
                nonzero = (hw >> 11) & 1
                if (r[rn] != 0) == bool(nonzero):
                    self.branch(self.pc_read() + imm)
            elif sub == 2:
                op, rm, rd = (hw >> 6) & 3, (hw >> 3) & 7, hw & 7
                v = r[rm]
                r[rd] = (_sx(v & 0xFFFF, 16) & M32, _sx(v & 0xFF, 8) & M32, v & 0xFFFF, v & 0xFF)[op]
            elif sub in (4, 5):                                   # PUSH
                regs = [i for i in range(8) if hw & (1 << i)] + ([14] if hw & 0x100 else [])
                a = r[13] - 4 * len(regs)
                r[13] = a & M32
                for i in regs:
                    self.wr32(a, r[i])
                    a += 4
            elif sub in (12, 13):                                 # POP
                regs = [i for i in range(8) if hw & (1 << i)] + ([15] if hw & 0x100 else [])
                a = r[13]
                for i in regs:
                    v = self.rd32(a)
                    a += 4
                    if i == 15:
                        self.branch_x(v)
                    else:
                        r[i] = v
                r[13] = a & M32
            elif sub == 15:                                       # IT / hints
                mask, first = hw & 15, (hw >> 4) & 15
                if mask == 0:
                    return                                        # NOP and friends
                conds = [first]
                m = mask
                while (m & 7) != 0:                               # bits above the terminating 1
                    conds.append((first & 14) | ((m >> 3) & 1))
                    m = (m << 1) & 15
                self.it = conds
            elif sub == 10:
                op, rm, rd = (hw >> 6) & 3, (hw >> 3) & 7, hw & 7
                v = r[rm]
                if op == 0:
                    r[rd] = struct.unpack("<I", struct.pack(">I", v))[0]
                else:
                    raise EmuError("REV16/REVSH")
            else:
                raise EmuError("misc16 %04x" % hw)
            return
        if (hw >> 12) == 12:                                      # STM / LDM
            load, rn = (hw >> 11) & 1, (hw >> 8) & 7
            regs = [i for i in range(8) if hw & (1 << i)]
            a = r[rn]
            for i in regs:
                if load:
                    r[i] = self.rd32(a)
                else:
                    self.wr32(a, r[i])
                a += 4
            if not load or rn not in regs:
                r[rn] = a & M32
            return
        if (hw >> 12) == 13:                                      # B<cond>
            c = (hw >> 8) & 15
            if c >= 14:
                raise EmuError("UDF/SVC")
            if self.cond(c):
                self.branch(self.pc_read() + _sx(hw & 0xFF, 8) * 2)
            return
        if (hw >> 11) == 0x1C:                                    # B
            self.branch(self.pc_read() + _sx(hw & 0x7FF, 11) * 2)
            return
        raise EmuError("thumb16 %04x at %08x" % (hw, self.r[15]))

    def branch_x(self, target):
        if (target & ~0xFF) == (self.RETURN_MAGIC & ~0xFF):
            self.next_pc = self.RETURN_MAGIC
        else:
            self.next_pc = target & ~1 & M32

    # ---- 32-bit ----
    def exec32(self, h, h2, in_it):
        r = self.r
        op1 = (h >> 11) & 3
        if (h & 0xEC00) == 0xEC00:                                # coprocessor space (VFP)
            return self.exec_vfp(h, h2)
        if op1 == 1:
            if (h & 0x0600) == 0:                                 # LDM/STM, LDRD/STRD, TBB
                if h & 0x0040:                                    # dual / exclusive / table branch
                    return self.ldst_dual(h, h2)
                return self.ldst_multiple(h, h2)
            if (h & 0x0600) == 0x0200:                            # data processing (shifted register)
                op, setf, rn = (h >> 5) & 15, (h >> 4) & 1, h & 15
                rd, rm = (h2 >> 8) & 15, h2 & 15
                imm5 = ((h2 >> 12) & 7) << 2 | ((h2 >> 6) & 3)
                typ, amt = self.decode_imm_shift((h2 >> 4) & 3, imm5)
                op2, cy = self.shift_c(self.reg(rm), typ, amt, self.c)
                rn_val = None if (rn == 15 and op in (2, 3)) else self.reg(rn)
                res = self.dp_op(op, rn_val, op2, cy, setf, rd)
                if res is not None:
                    r[rd] = res
                return
        if op1 == 2:
            if (h2 & 0x8000) == 0:
                if (h & 0x0200) == 0:                             # modified immediate
                    op, setf, rn, rd = (h >> 5) & 15, (h >> 4) & 1, h & 15, (h2 >> 8) & 15
                    imm12 = ((h >> 10) & 1) << 11 | ((h2 >> 12) & 7) << 8 | (h2 & 0xFF)
                    op2, cy = thumb_expand_imm(imm12, self.c)
                    rn_val = None if (rn == 15 and op in (2, 3)) else self.reg(rn)
                    res = self.dp_op(op, rn_val, op2, cy, setf, rd)
                    if res is not None:
                        r[rd] = res
                    return
                op, rn, rd = (h >> 4) & 31, h & 15, (h2 >> 8) & 15  # plain binary immediate
                imm12 = ((h >> 10) & 1) << 11 | ((h2 >> 12) & 7) << 8 | (h2 & 0xFF)
                if op == 0:                                       # ADDW / ADR
                    base = (self.pc_read() & ~3) if rn == 15 else r[rn]
                    r[rd] = (base + imm12) & M32
                elif op == 10:                                    # SUBW
                    base = (self.pc_read() & ~3) if rn == 15 else r[rn]
                    r[rd] = (base - imm12) & M32
                elif op == 4:                                     # MOVW
                    r[rd] = (rn << 12) | imm12
                elif op == 12:                                    # MOVT
                    r[rd] = (r[rd] & 0xFFFF) | (((rn << 12) | imm12) << 16)
                elif op in (20, 28):                              # SBFX / UBFX
                    lsb = ((h2 >> 12) & 7) << 2 | ((h2 >> 6) & 3)
                    width = (h2 & 31) + 1
                    v = (r[rn] >> lsb) & ((1 << width) - 1)
                    r[rd] = v if op == 28 else (_sx(v, width) & M32)
                elif op == 22:                                    # BFI / BFC
                    lsb = ((h2 >> 12) & 7) << 2 | ((h2 >> 6) & 3)
                    msb = h2 & 31
                    width = msb - lsb + 1
                    mask = ((1 << width) - 1) << lsb
                    src = 0 if rn == 15 else r[rn]
                    r[rd] = (r[rd] & ~mask & M32) | ((src << lsb) & mask)
                else:
                    raise EmuError("plain imm op %d" % op)
                return
            # branches and misc control
            s = (h >> 10) & 1
            j1, j2 = (h2 >> 13) & 1, (h2 >> 11) & 1
            if (h2 & 0x5000) == 0x0000:                           # conditional branch
                c = (h >> 6) & 15
                if c >= 14:
                    return                                        # hints / barriers
                imm = _sx((s << 20) | (j2 << 19) | (j1 << 18) | ((h & 0x3F) << 12) | ((h2 & 0x7FF) << 1), 21)
                if self.cond(c):
                    self.branch(self.pc_read() + imm)
                return
            i1, i2 = 1 ^ (j1 ^ s), 1 ^ (j2 ^ s)
            imm = _sx((s << 24) | (i1 << 23) | (i2 << 22) | ((h & 0x3FF) << 12) | ((h2 & 0x7FF) << 1), 25)
            if h2 & 0x4000:                                       # BL
                r[14] = (self.next_pc | 1) & M32
            self.branch_x(self.pc_read() + imm)
            return
        if op1 == 3:
            op2 = (h >> 4) & 0x7F
            if (op2 & 0x71) == 0x00:                              # store single data item
                return self.ldst_single(h, h2, load=False)
            if (op2 & 0x61) == 0x01:                              # loads (byte / half / word)
                return self.ldst_single(h, h2, load=True)
            if (op2 & 0x70) == 0x20:                              # data processing (register)
                if (h2 & 0x00F0) == 0 and (h2 & 0xF000) == 0xF000:   # LSL/LSR/ASR/ROR .W
                    typ, setf, rn, rd, rm = (h >> 5) & 3, (h >> 4) & 1, h & 15, (h2 >> 8) & 15, h2 & 15
                    res, cy = self.shift_c(r[rn], typ, r[rm] & 0xFF, self.c)
                    r[rd] = res
                    if setf:
                        self.nz(res, cy)
                    return
                if (h2 & 0x0080) and (h2 & 0xF000) == 0xF000 and (h & 0x0080) == 0:   # extend
                    op, rn, rd, rm = (h >> 4) & 7, h & 15, (h2 >> 8) & 15, h2 & 15
                    rot = ((h2 >> 4) & 3) * 8
                    v = ((r[rm] >> rot) | (r[rm] << (32 - rot))) & M32 if rot else r[rm]
                    ext = {0: _sx(v & 0xFFFF, 16) & M32, 1: v & 0xFFFF, 4: _sx(v & 0xFF, 8) & M32, 5: v & 0xFF}.get(op)
                    if ext is None:
                        raise EmuError("extend op %d" % op)
                    r[rd] = ext if rn == 15 else (r[rn] + ext) & M32
                    return
                raise EmuError("dp-reg %04x %04x" % (h, h2))
            if (op2 & 0x78) == 0x30:                              # multiply
                op, rn, ra, rd, rm = (h >> 4) & 7, h & 15, (h2 >> 12) & 15, (h2 >> 8) & 15, h2 & 15
                op2 = (h2 >> 4) & 3
                if op == 0 and op2 == 0:
                    r[rd] = (r[rn] * r[rm] + (0 if ra == 15 else r[ra])) & M32
                elif op == 0 and op2 == 1:
                    r[rd] = (r[ra] - r[rn] * r[rm]) & M32
                else:
                    raise EmuError("multiply %04x %04x" % (h, h2))
                return
            if (op2 & 0x78) == 0x38:                              # long multiply / divide
                op, rn, rlo, rhi, rm = (h >> 4) & 7, h & 15, (h2 >> 12) & 15, (h2 >> 8) & 15, h2 & 15
                if op == 2 and (h2 & 0xF0) == 0:                  # UMULL
                    p = r[rn] * r[rm]
                    r[rlo], r[rhi] = p & M32, (p >> 32) & M32
                elif op == 0 and (h2 & 0xF0) == 0:                # SMULL
                    p = _sx(r[rn], 32) * _sx(r[rm], 32)
                    r[rlo], r[rhi] = p & M32, (p >> 32) & M32
                elif op == 3 and (h2 & 0xF0) == 0xF0:             # UDIV
                    r[rhi] = (r[rn] // r[rm]) & M32 if r[rm] else 0
                elif op == 1 and (h2 & 0xF0) == 0xF0:             # SDIV
                    a, b = _sx(r[rn], 32), _sx(r[rm], 32)
                    q = 0 if b == 0 else abs(a) // abs(b) * (1 if (a < 0) == (b < 0) else -1)
                    r[rhi] = q & M32
                else:
                    raise EmuError("long mul %04x %04x" % (h, h2))
                return
        raise EmuError("thumb32 %04x %04x at %08x" % (h, h2, self.r[15]))

    def ldst_multiple(self, h, h2):
        r = self.r
        op, w, load, rn = (h >> 7) & 3, (h >> 5) & 1, (h >> 4) & 1, h & 15
        regs = [i for i in range(16) if h2 & (1 << i)]
        if op == 1:                                               # increment after
            a = r[rn]
            end = a + 4 * len(regs)
        elif op == 2:                                             # decrement before
            a = r[rn] - 4 * len(regs)
            end = a
        else:
            raise EmuError("ldm/stm mode")
        for i in regs:
            if load:
                v = self.rd32(a)
                if i == 15:
                    self.branch_x(v)
                else:
                    r[i] = v
            else:
                self.wr32(a, r[i])
            a += 4
        if w and not (load and rn in regs):
            r[rn] = end & M32

    def ldst_dual(self, h, h2):
        r = self.r
        p, u, w, load, rn = (h >> 8) & 1, (h >> 7) & 1, (h >> 5) & 1, (h >> 4) & 1, h & 15
        if not p and not w:
            if (h & 0x00F0) == 0x00D0 and (h2 & 0xFFE0) == 0xF000:    # TBB / TBH
                rm = h2 & 15
                if h2 & 0x10:
                    off = self.rd16((self.reg(rn) + 2 * r[rm]) & M32)
                else:
                    off = self.rd8((self.reg(rn) + r[rm]) & M32)
                self.branch(self.pc_read() + 2 * off)
                return
            raise EmuError("exclusive %04x %04x" % (h, h2))
        rt, rt2, imm = (h2 >> 12) & 15, (h2 >> 8) & 15, (h2 & 0xFF) * 4
        base = (self.pc_read() & ~3) if rn == 15 else r[rn]
        off = (base + imm if u else base - imm) & M32
        a = off if p else base
        if load:
            r[rt], r[rt2] = self.rd32(a), self.rd32(a + 4)
        else:
            self.wr32(a, r[rt])
            self.wr32(a + 4, r[rt2])
        if w:
            r[rn] = off

    def ldst_single(self, h, h2, load):
        r = self.r
        size = (h >> 5) & 3                                       # 0 byte, 1 half, 2 word
        signed = (h >> 8) & 1
        rn, rt = h & 15, (h2 >> 12) & 15
        if rn == 15:                                              # literal
            if not load:
                raise EmuError("store literal")
            u = (h >> 7) & 1
            base = self.pc_read() & ~3
            a = base + (h2 & 0xFFF) if u else base - (h2 & 0xFFF)
            wb = None
        elif h & 0x0080:                                          # imm12, positive
            a = (r[rn] + (h2 & 0xFFF)) & M32
            wb = None
        elif h2 & 0x0800:                                         # imm8 with P/U/W
            p, u, w = (h2 >> 10) & 1, (h2 >> 9) & 1, (h2 >> 8) & 1
            imm = h2 & 0xFF
            off = (r[rn] + imm if u else r[rn] - imm) & M32
            a = off if p else r[rn]
            wb = off if w else None
        elif (h2 & 0x0FC0) == 0:                                  # register offset, LSL #imm2
            a = (r[rn] + (r[h2 & 15] << ((h2 >> 4) & 3))) & M32
            wb = None
        else:
            raise EmuError("ldst single %04x %04x" % (h, h2))
        if load:
            if size == 2:
                v = self.rd32(a)
            elif size == 1:
                v = self.rd16(a)
                if signed:
                    v = _sx(v, 16) & M32
            else:
                v = self.rd8(a)
                if signed:
                    v = _sx(v, 8) & M32
            if wb is not None:
                r[rn] = wb
            if rt == 15:
                if size != 2:
                    return                                        # PLD and friends
                self.branch_x(v)
            else:
                r[rt] = v
        else:
            if size == 2:
                self.wr32(a, r[rt])
            elif size == 1:
                self.wr16(a, r[rt])
            else:
                self.wr8(a, r[rt])
            if wb is not None:
                r[rn] = wb

    # ---- VFP (single precision only) ----
    def exec_vfp(self, h, h2):
        r, s = self.r, self.s
        if (h2 & 0x0E00) != 0x0A00:
            raise EmuError("coprocessor %04x %04x" % (h, h2))
        dbl = (h2 >> 8) & 1
        if (h & 0x0F00) == 0x0E00 and (h2 & 0x0010) == 0:         # data processing
            if dbl:
                raise EmuError("double-precision VFP")
            opc1 = (h >> 4) & 0xB
            d = (((h2 >> 12) & 15) << 1) | ((h >> 6) & 1)
            nidx = ((h & 15) << 1) | ((h2 >> 7) & 1)
            m = ((h2 & 15) << 1) | ((h2 >> 5) & 1)
            op = (h2 >> 6) & 1
            if opc1 != 0xB:
                a, b = _f(s[nidx]), _f(s[m])
                acc = _f(s[d])
                if opc1 == 0:                                     # VMLA / VMLS (two roundings)
                    p = np.float32(a * b)
                    res = np.float32(acc - p) if op else np.float32(acc + p)
                elif opc1 == 1:                                   # VNMLS / VNMLA
                    p = np.float32(a * b)
                    res = np.float32(-acc - p) if op else np.float32(-acc + p)
                elif opc1 == 2:
                    res = np.float32(a * b)
                    if op:
                        res = -res
                elif opc1 == 3:
                    res = np.float32(a - b) if op else np.float32(a + b)
                elif opc1 == 8:
                    res = np.float32(a / b)
                elif opc1 in (9, 10):                             # fused: VFNMS/VFNMA (9), VFMA/VFMS (10)
                    aa = -float(a) if op else float(a)
                    cc = float(acc)
                    if opc1 == 9:
                        aa, cc = -aa, -cc
                    res = _fma32(aa, float(b), cc)
                else:
                    raise EmuError("vfp opc1 %x" % opc1)
                s[d] = _u(res)
                return
            opc2 = h & 15
            if (h2 & 0x0040) == 0:                                # VMOV immediate
                imm8 = (opc2 << 4) | (h2 & 15)
                sign, bexp = imm8 >> 7, (imm8 >> 6) & 1
                exp = ((bexp ^ 1) << 7) | ((0x1F if bexp else 0) << 2) | ((imm8 >> 4) & 3)
                s[d] = (sign << 31) | (exp << 23) | ((imm8 & 15) << 19)
                return
            hi = (h2 >> 7) & 1
            if opc2 == 0:
                s[d] = s[m] & 0x7FFFFFFF if hi else s[m]          # VABS / VMOV
            elif opc2 == 1:
                s[d] = _u(np.sqrt(_f(s[m]))) if hi else s[m] ^ 0x80000000
            elif opc2 in (4, 5):                                  # VCMP{E}
                a = _f(s[d])
                b = np.float32(0) if opc2 == 5 else _f(s[m])
                if np.isnan(a) or np.isnan(b):
                    self.fn, self.fz, self.fc, self.fv = 0, 0, 1, 1
                elif a == b:
                    self.fn, self.fz, self.fc, self.fv = 0, 1, 1, 0
                elif a < b:
                    self.fn, self.fz, self.fc, self.fv = 1, 0, 0, 0
                else:
                    self.fn, self.fz, self.fc, self.fv = 0, 0, 1, 0
            elif opc2 == 8:                                       # VCVT.F32.{U32,S32}
                v = s[m]
                s[d] = _u(np.float32(_sx(v, 32) if hi else v))
            elif opc2 in (12, 13):                                # VCVT{R}.{U32,S32}.F32
                if not hi:
                    raise EmuError("VCVTR")
                a = float(_f(s[m]))
                if np.isnan(a):
                    v = 0
                else:
                    t = int(a) if abs(a) < 1e30 else (1 << 40) * (1 if a > 0 else -1)
                    v = max(-(1 << 31), min((1 << 31) - 1, t)) if opc2 == 13 else max(0, min(M32, t))
                s[d] = v & M32
            else:
                raise EmuError("vfp ext opc2 %x" % opc2)
            return
        if (h & 0x0FE0) == 0x0E00 and (h2 & 0x0F7F) == 0x0A10:    # VMOV core <-> single
            idx = ((h & 15) << 1) | ((h2 >> 7) & 1)
            rt = (h2 >> 12) & 15
            if h & 0x10:
                r[rt] = s[idx]
            else:
                s[idx] = r[rt]
            return
        if (h & 0x0FFF) == 0x0EF1 and (h2 & 0x0FFF) == 0x0A10:    # VMRS
            rt = (h2 >> 12) & 15
            if rt == 15:
                self.n, self.z, self.c, self.v = self.fn, self.fz, self.fc, self.fv
            else:
                r[rt] = (self.fn << 31) | (self.fz << 30) | (self.fc << 29) | (self.fv << 28)
            return
        if (h & 0x0FF0) == 0x0EE0 and (h2 & 0x0FFF) == 0x0A10:    # VMSR: rounding mode stays RN
            return
        if (h & 0x0FE0) == 0x0C40:                                # VMOV two core regs <-> two singles / double
            rt, rt2 = (h2 >> 12) & 15, h & 15
            m = ((h2 & 15) << 1) | ((h2 >> 5) & 1)
            if dbl:
                m = (h2 & 15) << 1
            if h & 0x10:
                r[rt], r[rt2] = s[m], s[m + 1]
            else:
                s[m], s[m + 1] = r[rt], r[rt2]
            return
        if (h & 0x0E00) == 0x0C00:                                # VLDR/VSTR/VLDM/VSTM/VPUSH/VPOP
            p, u, w, load, rn = (h >> 8) & 1, (h >> 7) & 1, (h >> 5) & 1, (h >> 4) & 1, h & 15
            imm = (h2 & 0xFF) * 4
            vd = (h2 >> 12) & 15
            d = (vd << 1) if dbl else ((vd << 1) | ((h >> 6) & 1))
            if p and not w:                                       # VLDR / VSTR
                base = (self.pc_read() & ~3) if rn == 15 else r[rn]
                a = (base + imm if u else base - imm) & M32
                cnt = 2 if dbl else 1
                for i in range(cnt):
                    if load:
                        s[d + i] = self.rd32(a + 4 * i)
                    else:
                        self.wr32(a + 4 * i, s[d + i])
                return
            words = h2 & 0xFF
            if p == u:
                raise EmuError("vldm mode")
            a = r[rn] if u else (r[rn] - imm) & M32
            for i in range(words):
                if load:
                    s[d + i] = self.rd32(a + 4 * i)
                else:
                    self.wr32(a + 4 * i, s[d + i])
            if w:
                r[rn] = (r[rn] + imm if u else r[rn] - imm) & M32
            return
        raise EmuError("vfp %04x %04x at %08x" % (h, h2, self.r[15]))


def _fma32(a, b, c):
    """fused multiply-add rounded once to float32 (exact rational arithmetic; only used if the archive has VFMA)"""
    from fractions import Fraction
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        return np.float32(a * b + c)
    exact = Fraction(a) * Fraction(b) + Fraction(c)
    if exact == 0:
        return np.float32(a * b + c)
    # round the exact value to float32: float64 conversion of a Fraction is correctly rounded; double rounding is
    # avoided by checking the halfway case
    f64 = float(exact)
    f32 = np.float32(f64)
    lo = np.nextafter(f32, np.float32(-np.inf))
    hi = np.nextafter(f32, np.float32(np.inf))
    best = min((lo, f32, hi), key=lambda v: (abs(Fraction(float(v)) - exact), struct.unpack("<I", np.float32(v).tobytes())[0] & 1))
    return np.float32(best)


# ---- a minimal static linker for relocatable ELF32 members of an `ar` archive --------------------------
def read_archive(path):
    data = open(path, "rb").read()
    assert data[:8] == b"!<arch>\n"
    off, members, longnames = 8, {}, b""
    while off + 60 <= len(data):
        hdr = data[off:off + 60]
        name = hdr[:16].decode().rstrip()
        size = int(hdr[48:58].decode().strip())
        body = data[off + 60:off + 60 + size]
        if name == "//":
            longnames = body
        elif name.startswith("/") and name[1:].isdigit():
            i = int(name[1:])
            name = longnames[i:longnames.index(b"/", i)].decode()
            members[name] = body
        elif name not in ("/", "/SYM64/"):
            members[name.rstrip("/")] = body
        off += 60 + size + (size & 1)
    return members


class Linker:
    def __init__(self, cpu, base=0x10000):
        self.cpu = cpu
        self.cursor = base
        self.symbols = {}
        self.pending = []
        self.allow_undefined = True                               # data of other number formats (q15/q31 tables) is not loaded
        self.undefined = set()

    def add_hook(self, name, fn):
        a = self.alloc(4, 4)
        self.cpu.hooks[a] = fn
        self.symbols[name] = a

    def alloc(self, size, align=8):
        self.cursor = (self.cursor + align - 1) & ~(align - 1)
        a = self.cursor
        self.cursor += size
        return a

    def add_object(self, blob):
        e_shoff = struct.unpack_from("<I", blob, 0x20)[0]
        shentsize, shnum, shstrndx = struct.unpack_from("<HHH", blob, 0x2E)
        sh = [struct.unpack_from("<10I", blob, e_shoff + i * shentsize) for i in range(shnum)]
        addr = {}
        for i, (nm, typ, flags, _, off, size, link, info, align, entsize) in enumerate(sh):
            if flags & 2 and typ in (1, 8):                       # SHF_ALLOC progbits / nobits
                a = self.alloc(size, max(align, 4))
                if typ == 1:
                    self.cpu.mem[a:a + size] = blob[off:off + size]
                addr[i] = a
        symtab = next(s for s in sh if s[1] == 2)
        strtab = sh[symtab[6]]
        nsym = symtab[5] // 16
        syms = []
        for i in range(nsym):
            st_name, st_value, st_size, st_info, st_other, st_shndx = struct.unpack_from("<IIIBBH", blob, symtab[4] + 16 * i)
            name = blob[strtab[4] + st_name:blob.index(b"\0", strtab[4] + st_name)].decode()
            syms.append((name, st_value, st_info, st_shndx))
            if st_shndx in addr and (st_info >> 4) in (1, 2) and name:
                self.symbols[name] = addr[st_shndx] + st_value
        for s in sh:
            if s[1] == 9 and s[7] in addr:                        # SHT_REL applying to a loaded section
                for k in range(s[5] // 8):
                    r_off, r_info = struct.unpack_from("<II", blob, s[4] + 8 * k)
                    self.pending.append((addr[s[7]] + r_off, r_info & 0xFF, syms[r_info >> 8], addr))

    def resolve(self):
        cpu = self.cpu
        for place, rtype, (name, value, info, shndx), addr in self.pending:
            if shndx in addr:
                sval = addr[shndx] + value
            elif name in self.symbols:
                sval = self.symbols[name]
            elif self.allow_undefined:
                sval = 0x00DEAD00                                 # never mapped to anything this tool runs
                self.undefined.add(name)
            else:
                raise EmuError("undefined symbol %s" % name)
            is_func = (info & 15) == 2
            if rtype in (2, 38):                                  # R_ARM_ABS32 / TARGET1
                cpu.wr32(place, (cpu.rd32(place) + sval) | (1 if is_func else 0))
            elif rtype in (10, 30):                               # R_ARM_THM_CALL / JUMP24
                h, h2 = cpu.rd16(place), cpu.rd16(place + 2)
                sbit = (h >> 10) & 1
                j1, j2 = (h2 >> 13) & 1, (h2 >> 11) & 1
                i1, i2 = 1 ^ (j1 ^ sbit), 1 ^ (j2 ^ sbit)
                addend = _sx((sbit << 24) | (i1 << 23) | (i2 << 22) | ((h & 0x3FF) << 12) | ((h2 & 0x7FF) << 1), 25)
                off = (sval & ~1) + addend - place
                sbit = (off >> 24) & 1
                i1, i2 = (off >> 23) & 1, (off >> 22) & 1
                j1, j2 = (1 ^ i1) ^ sbit, (1 ^ i2) ^ sbit
                cpu.wr16(place, (h & 0xF800) | (sbit << 10) | ((off >> 12) & 0x3FF))
                cpu.wr16(place + 2, (h2 & 0xD000) | (j1 << 13) | (j2 << 11) | ((off >> 1) & 0x7FF))
            elif rtype in (47, 48):                               # R_ARM_THM_MOVW_ABS_NC / MOVT_ABS
                h, h2 = cpu.rd16(place), cpu.rd16(place + 2)
                addend = _sx(((h & 15) << 12) | (((h >> 10) & 1) << 11) | (((h2 >> 12) & 7) << 8) | (h2 & 0xFF), 16)
                v = (sval + addend) | (1 if is_func and rtype == 47 else 0)
                if rtype == 48:
                    v >>= 16
                v &= 0xFFFF
                cpu.wr16(place, (h & 0xFBF0) | (v >> 12) | (((v >> 11) & 1) << 10))
                cpu.wr16(place + 2, (h2 & 0x8F00) | (((v >> 8) & 7) << 12) | (v & 0xFF))
            elif rtype == 40:                                     # R_ARM_V4BX
                pass
            elif rtype == 102:                                    # R_ARM_THM_JUMP11
                h = cpu.rd16(place)
                off = (sval & ~1) + _sx(h & 0x7FF, 11) * 2 - place
                cpu.wr16(place, (h & 0xF800) | ((off >> 1) & 0x7FF))
            elif rtype == 103:                                    # R_ARM_THM_JUMP8
                h = cpu.rd16(place)
                off = (sval & ~1) + _sx(h & 0xFF, 8) * 2 - place
                cpu.wr16(place, (h & 0xFF00) | ((off >> 1) & 0xFF))
            else:
                raise EmuError("relocation type %d for %s" % (rtype, name))
        self.pending = []
