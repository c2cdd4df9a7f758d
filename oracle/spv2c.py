"""oracle/spv2c.py — TEST INFRASTRUCTURE ONLY (never imported by the product).

Translates one of the reference's SHIPPED SPIR-V modules (/root/reference/compiled-shaders/**.spv, built by
rust-gpu from shader/src/*.rs + glam-pbr + shared-structs) into one plain-C function, so that the reference's own
compiled code can be executed on the CPU:  `python oracle/spv2c.py in.spv out.c`.  oracle/build_ref.py drives it
and compiles the result into oracle/_ref/ (git-ignored; generated from the binaries where they lie, nothing of
the reference is copied into the repository).

What the translation keeps: every arithmetic instruction, in the module's order, one C statement per SPIR-V
instruction, IEEE binary32, no contraction (compile with -ffp-contract=off).  GLSL.std.450 calls map onto glibc
(sqrtf is exact; powf/expf/logf/log2f/sinf/cosf are the same functions the C oracle calls, so the two agree
bit for bit wherever the restatement follows the module's operation order).  What the environment supplies
through `spv_ctx` callbacks, because fixed-function hardware did it in the reference: image sampling
(OpImageSample*Lod) and screen-space derivatives (OpDPdx/OpDPdy).  Atomics run sequentially (one invocation at
a time), so list ORDER is the invocation order, not the reference's race order.

Supported: the opcode set rust-gpu emitted for these modules (a single fully inlined function, structured
control flow, logical addressing, no matrices, no OpFunctionCall).  Anything else raises.
"""
import struct
import sys

# storage classes
SC_UNIFORM_CONSTANT, SC_INPUT, SC_UNIFORM, SC_OUTPUT, SC_PRIVATE, SC_FUNCTION, SC_PUSH, SC_STORAGE = 0, 1, 2, 3, 6, 7, 9, 12
MEMORY_CLASSES = (SC_UNIFORM, SC_PUSH, SC_STORAGE)
# decorations
D_BLOCK, D_BUFFER_BLOCK, D_ARRAY_STRIDE, D_BUILTIN, D_LOCATION, D_BINDING, D_SET, D_OFFSET = 2, 3, 6, 11, 30, 33, 34, 35


class T:
    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)


class Module:
    def __init__(self, words):
        assert words[0] == 0x07230203, "not SPIR-V"
        self.bound = words[3]
        self.types, self.consts, self.deco, self.mdeco = {}, {}, {}, {}
        self.globals, self.body = [], []
        self.entry_name, self.exec_model, self.local_size = None, None, (1, 1, 1)
        self.ext_sets = {}
        i = 5
        in_fn = False
        while i < len(words):
            wc, op = words[i] >> 16, words[i] & 0xFFFF
            a = list(words[i + 1:i + wc])
            i += wc
            if in_fn:
                if op == 56:
                    in_fn = False
                else:
                    self.body.append((op, a))
                continue
            if op == 54:
                in_fn = True
            elif op == 11:
                self.ext_sets[a[0]] = self._str(a[1:])
            elif op == 15:
                self.exec_model, self.entry_id = a[0], a[1]
                self.entry_name = self._str(a[2:])
            elif op == 16 and a[1] == 17:
                self.local_size = tuple(a[2:5])
            elif op == 71:
                self.deco.setdefault(a[0], {})[a[1]] = a[2:]
            elif op == 72:
                self.mdeco.setdefault((a[0], a[1]), {})[a[2]] = a[3:]
            elif op == 19:
                self.types[a[0]] = T("void")
            elif op == 20:
                self.types[a[0]] = T("bool")
            elif op == 21:
                self.types[a[0]] = T("int", width=a[1], signed=a[2])
            elif op == 22:
                self.types[a[0]] = T("float", width=a[1])
            elif op == 23:
                self.types[a[0]] = T("vector", elem=a[1], n=a[2])
            elif op == 4472:
                self.types[a[0]] = T("rayquery")      # OpTypeRayQueryKHR
            elif op == 5341:
                self.types[a[0]] = T("accel")         # OpTypeAccelerationStructureKHR
            elif op == 25:
                self.types[a[0]] = T("image")
            elif op == 26:
                self.types[a[0]] = T("sampler")
            elif op == 27:
                self.types[a[0]] = T("sampled_image")
            elif op == 28:
                self.types[a[0]] = T("array", elem=a[1], length_id=a[2])
            elif op == 29:
                self.types[a[0]] = T("rtarray", elem=a[1])
            elif op == 30:
                self.types[a[0]] = T("struct", members=a[1:])
            elif op == 32:
                self.types[a[0]] = T("pointer", sc=a[1], pointee=a[2])
            elif op == 33:
                self.types[a[0]] = T("function")
            elif op in (41, 42):
                self.consts[a[1]] = (a[0], 1 if op == 41 else 0)
            elif op == 43:
                self.consts[a[1]] = (a[0], a[2] | (a[3] << 32) if len(a) > 3 else a[2])
            elif op == 44:
                self.consts[a[1]] = (a[0], tuple(a[2:]))
            elif op == 59:
                self.globals.append((a[1], a[0], a[2]))
            elif op in (10, 14, 17, 16, 5, 6, 3, 4, 7, 8):
                pass
            else:
                raise NotImplementedError(f"module-level opcode {op}")

    @staticmethod
    def _str(ws):
        return b"".join(struct.pack("<I", x) for x in ws).split(b"\0")[0].decode()

    def array_len(self, t):
        return self.consts[t.length_id][1]


class Emitter:
    def __init__(self, m, symbol):
        self.m, self.symbol = m, symbol
        self.id_type = {}  # SSA id -> type id
        self.decl, self.code, self.helpers, self.helper_order = [], [], {}, []
        self.typedefs, self.typedef_done = [], set()
        self.ptr_info = {}  # id -> ('mem', pointee type id) | ('val', ...) | ('img', ...)

    # ---- C types -------------------------------------------------------------------------------------
    def ctype(self, tid):
        t = self.m.types[tid]
        k = t.kind
        if k == "bool":
            return "int"
        if k == "int":
            assert t.width in (32, 64)
            if t.width == 64:
                return "int64_t" if t.signed else "uint64_t"
            return "int32_t" if t.signed else "uint32_t"
        if k == "float":
            assert t.width == 32
            return "float"
        if k == "rayquery":
            return "spv_rayq"
        if k == "accel":
            return "uint64_t"
        if k in ("image", "sampler"):
            return "spv_handle"
        if k == "sampled_image":
            return "spv_sampled"
        if k == "pointer":
            if t.sc in MEMORY_CLASSES:
                return "uint8_t*"
            if t.sc == SC_UNIFORM_CONSTANT:
                return "spv_handle"
            return self.ctype(t.pointee) + "*"
        name = f"T{tid}"
        if tid not in self.typedef_done:
            self.typedef_done.add(tid)
            if k == "vector":
                self.typedefs.append(f"typedef struct {{ {self.ctype(t.elem)} v[{t.n}]; }} {name};")
            elif k == "array":
                self.typedefs.append(f"typedef struct {{ {self.ctype(t.elem)} a[{self.m.array_len(t)}]; }} {name};")
            elif k == "struct":
                fields = " ".join(f"{self.ctype(mt)} m{j};" for j, mt in enumerate(t.members))
                self.typedefs.append(f"typedef struct {{ {fields} }} {name};")
            else:
                raise NotImplementedError(f"value of type kind {k}")
        return name

    # ---- memory <-> value with the module's explicit layout ------------------------------------------
    def loader(self, tid, store=False):
        key = ("st" if store else "ld", tid)
        if key in self.helpers:
            return self.helpers[key][0]
        t = self.m.types[tid]
        ct = self.ctype(tid)
        fn = f"{'st' if store else 'ld'}_{tid}"
        self.helpers[key] = (fn, None)
        lines = []
        if t.kind in ("int", "float"):
            nb = t.width // 8
            body = f"memcpy(p, &x, {nb});" if store else f"{ct} x; memcpy(&x, p, {nb}); return x;"
        elif t.kind == "vector":
            es = 4
            sub = self.loader(t.elem, store)
            if store:
                body = " ".join(f"{sub}(p + {j * es}, x.v[{j}]);" for j in range(t.n))
            else:
                body = f"{ct} x; " + " ".join(f"x.v[{j}] = {sub}(p + {j * es});" for j in range(t.n)) + " return x;"
        elif t.kind == "array":
            stride = self.m.deco[tid][D_ARRAY_STRIDE][0]
            sub = self.loader(t.elem, store)
            n = self.m.array_len(t)
            if store:
                body = f"for (int j = 0; j < {n}; j++) {sub}(p + j * {stride}, x.a[j]);"
            else:
                body = f"{ct} x; for (int j = 0; j < {n}; j++) x.a[j] = {sub}(p + j * {stride}); return x;"
        elif t.kind == "struct":
            parts = []
            for j, mt in enumerate(t.members):
                off = self.m.mdeco[(tid, j)][D_OFFSET][0]
                sub = self.loader(mt, store)
                parts.append(f"{sub}(p + {off}, x.m{j});" if store else f"x.m{j} = {sub}(p + {off});")
            body = " ".join(parts) if store else f"{ct} x; " + " ".join(parts) + " return x;"
        else:
            raise NotImplementedError(f"load/store of {t.kind}")
        sig = f"static inline void {fn}(uint8_t* p, {ct} x)" if store else f"static inline {ct} {fn}(const uint8_t* p)"
        self.helpers[key] = (fn, f"{sig} {{ {body} }}")
        self.helper_order.append(key)
        return fn

    # ---- values --------------------------------------------------------------------------------------
    def v(self, i):
        return f"_{i}"

    def const_init(self, cid):
        tid, val = self.m.consts[cid]
        t = self.m.types[tid]
        if t.kind == "bool":
            return str(val)
        if t.kind == "int" and t.width == 64:
            return f"(int64_t)0x{val:016x}ull" if t.signed else f"0x{val:016x}ull"
        if t.kind == "int":
            return f"(int32_t)0x{val:08x}u" if t.signed else f"0x{val:08x}u"
        if t.kind == "float":
            return f"spv_bits2f(0x{val:08x}u)"
        if t.kind in ("vector", "array", "struct"):
            return "{" + ("{" if t.kind != "struct" else "") + ", ".join(self.const_init(c) for c in val) + \
                   ("}" if t.kind != "struct" else "") + "}"
        raise NotImplementedError(t.kind)

    def ncomp(self, tid):
        t = self.m.types[tid]
        return t.n if t.kind == "vector" else 0

    def elementwise(self, r, rt, fmt, *ops):
        """fmt uses {0},{1}.. for operand component expressions."""
        n = self.ncomp(rt)
        if n == 0:
            return f"{self.v(r)} = {fmt.format(*[self.v(o) for o in ops])};"
        out = []
        for j in range(n):
            exprs = []
            for o in ops:
                ot = self.m.types[self.id_type[o]]
                exprs.append(f"{self.v(o)}.v[{j}]" if ot.kind == "vector" else self.v(o))
            out.append(f"{self.v(r)}.v[{j}] = {fmt.format(*exprs)};")
        return " ".join(out)

    # ---- translation ---------------------------------------------------------------------------------
    def run(self):
        m = self.m
        for cid, (tid, _) in m.consts.items():
            self.id_type[cid] = tid
        pre = []
        for cid in m.consts:
            tid = m.consts[cid][0]
            pre.append(f"const {self.ctype(tid)} {self.v(cid)} = {self.const_init(cid)};")
        # global variables
        self.var_meta = {}
        for vid, ptid, sc in m.globals:
            self.id_type[vid] = ptid
            pt = m.types[ptid]
            d = m.deco.get(vid, {})
            if sc in (SC_UNIFORM, SC_STORAGE):
                s, b = d[D_SET][0], d[D_BINDING][0]
                pre.append(f"uint8_t* const {self.v(vid)} = ctx->buf[{s}][{b}].ptr;")
                self.var_meta[vid] = ("buf", s, b)
            elif sc == SC_PUSH:
                pre.append(f"uint8_t* const {self.v(vid)} = ctx->push;")
            elif sc == SC_UNIFORM_CONSTANT:
                s, b = d[D_SET][0], d[D_BINDING][0]
                pre.append(f"const spv_handle {self.v(vid)} = {{{s}, {b}, 0}};")
            elif sc in (SC_INPUT, SC_OUTPUT):
                ct = self.ctype(pt.pointee)
                if D_BUILTIN in d:
                    pre.append(f"{ct}* const {self.v(vid)} = ({ct}*)ctx->builtin[{d[D_BUILTIN][0]}];")
                else:
                    arr = "in_loc" if sc == SC_INPUT else "out_loc"
                    pre.append(f"{ct}* const {self.v(vid)} = ({ct}*)ctx->{arr}[{d[D_LOCATION][0]}];")
                    self.var_meta[vid] = ("loc", d[D_LOCATION][0])
            elif sc == SC_PRIVATE:
                ct = self.ctype(pt.pointee)
                pre.append(f"{ct} {self.v(vid)}_s; memset(&{self.v(vid)}_s, 0, sizeof {self.v(vid)}_s); {ct}* const {self.v(vid)} = &{self.v(vid)}_s;")
            else:
                raise NotImplementedError(f"global storage class {sc}")

        # pass 1: result types, phis per block, block order
        blocks, cur = [], None
        phis = {}  # block label -> [(result, type, [(value, parent)...])]
        for op, a in m.body:
            if op == 248:
                cur = a[0]
                blocks.append(cur)
                phis[cur] = []
            elif op == 245:
                phis[cur].append((a[1], a[0], list(zip(a[2::2], a[3::2]))))
                self.id_type[a[1]] = a[0]
            elif op in RESULT_OPS:
                self.id_type[a[1]] = a[0]
        self.load_src = {}  # load result -> pointer id (for derivative tracing)

        def edge(src, dst):
            ps = [(r, t, dict((p, v) for v, p in inc)) for r, t, inc in phis[dst]]
            if not ps:
                return f"goto L{dst};"
            s = "{ "
            for k, (r, t, inc) in enumerate(ps):
                s += f"{self.ctype(t)} t{k} = {self.v(inc[src])}; "
            for k, (r, t, inc) in enumerate(ps):
                s += f"{self.v(r)} = t{k}; "
            return s + f"goto L{dst}; }}"

        code = self.code
        cur = None
        for op, a in m.body:
            if op == 248:
                cur = a[0]
                code.append(f"L{cur}: ;")
                continue
            if op in (246, 247, 245):
                continue
            if op == 249:
                code.append(edge(cur, a[0]))
                continue
            if op == 250:
                code.append(f"if ({self.v(a[0])}) {edge(cur, a[1])} else {edge(cur, a[2])}")
                continue
            if op == 251:
                s = f"switch ((uint32_t){self.v(a[0])}) {{ "
                for lit, lab in zip(a[2::2], a[3::2]):
                    s += f"case {lit}u: {edge(cur, lab)} "
                s += f"default: {edge(cur, a[1])} }}"
                code.append(s)
                continue
            if op == 252:
                code.append("ctx->killed = 1; return;")
                continue
            if op == 253:
                code.append("return;")
                continue
            if op == 255:
                code.append("abort();")
                continue
            code.append(self.instr(op, a) + f"  /* op{op} */")

        decls = []
        for i, tid in sorted(self.id_type.items()):
            if i in m.consts or any(i == g[0] for g in m.globals):
                continue
            if m.types[tid].kind in ("void", "function"):
                continue
            decls.append(f"{self.ctype(tid)} {self.v(i)};")
        out = [PRELUDE]
        out += self.typedefs
        out += [self.helpers[k][1] for k in self.helper_order]
        out.append(f"void {self.symbol}(spv_ctx* ctx) {{")
        out += pre + decls + self.decl + code
        out.append("}")
        ls = m.local_size
        out.append(f"const uint32_t {self.symbol}_local_size[3] = {{{ls[0]}, {ls[1]}, {ls[2]}}};")
        out.append(f"const uint32_t {self.symbol}_execution_model = {m.exec_model};")
        return "\n".join(out) + "\n"

    def instr(self, op, a):
        m, v = self.m, self.v
        if op == 59:  # function-local variable
            ptid, r = a[0], a[1]
            ct = self.ctype(m.types[ptid].pointee)
            self.decl.append(f"{ct} {v(r)}_s; memset(&{v(r)}_s, 0, sizeof {v(r)}_s);")
            return f"{v(r)} = &{v(r)}_s;"
        if op == 61:
            rt, r, p = a[0], a[1], a[2]
            pt = m.types[self.id_type[p]]
            self.load_src[r] = p
            if pt.sc in MEMORY_CLASSES:
                return f"{v(r)} = {self.loader(rt)}({v(p)});"
            if pt.sc == SC_UNIFORM_CONSTANT:
                return f"{v(r)} = {v(p)};"
            return f"{v(r)} = *{v(p)};"
        if op == 62:
            p, x = a[0], a[1]
            pt = m.types[self.id_type[p]]
            if pt.sc in MEMORY_CLASSES:
                return f"{self.loader(pt.pointee, store=True)}({v(p)}, {v(x)});"
            return f"*{v(p)} = {v(x)};"
        if op == 65:
            rt, r, base, idx = a[0], a[1], a[2], a[3:]
            bt = m.types[self.id_type[base]]
            tid = bt.pointee
            if bt.sc in MEMORY_CLASSES:
                expr = v(base)
                for ix in idx:
                    t = m.types[tid]
                    if t.kind == "struct":
                        k = m.consts[ix][1]
                        expr += f" + {m.mdeco[(tid, k)][D_OFFSET][0]}"
                        tid = t.members[k]
                    elif t.kind in ("array", "rtarray"):
                        expr += f" + (size_t)(uint32_t){v(ix)} * {m.deco[tid][D_ARRAY_STRIDE][0]}"
                        tid = t.elem
                    elif t.kind == "vector":
                        expr += f" + (size_t)(uint32_t){v(ix)} * 4"
                        tid = t.elem
                    else:
                        raise NotImplementedError(t.kind)
                return f"{v(r)} = {expr};"
            if bt.sc == SC_UNIFORM_CONSTANT:
                assert len(idx) == 1
                return f"{v(r)} = {v(base)}; {v(r)}.index = (uint32_t){v(idx[0])};"
            expr = f"(*{v(base)})"
            for ix in idx:
                t = m.types[tid]
                if t.kind == "struct":
                    k = m.consts[ix][1]
                    expr += f".m{k}"
                    tid = t.members[k]
                elif t.kind == "array":
                    expr += f".a[(uint32_t){v(ix)}]"
                    tid = t.elem
                elif t.kind == "vector":
                    expr += f".v[(uint32_t){v(ix)}]"
                    tid = t.elem
                else:
                    raise NotImplementedError(t.kind)
            return f"{v(r)} = &{expr};"
        if op == 68:
            r, sp, member = a[1], a[2], a[3]
            stid = m.types[self.id_type[sp]].pointee
            off = m.mdeco[(stid, member)][D_OFFSET][0]
            stride = m.deco[m.types[stid].members[member]][D_ARRAY_STRIDE][0]
            _, s, b = self.var_meta[sp]
            return f"{v(r)} = (uint32_t)((ctx->buf[{s}][{b}].size - {off}) / {stride});"
        if op == 80:
            rt, r, parts = a[0], a[1], a[2:]
            t = m.types[rt]
            if t.kind == "vector":
                out, j = [], 0
                for p in parts:
                    pn = self.ncomp(self.id_type[p])
                    if pn == 0:
                        out.append(f"{v(r)}.v[{j}] = {v(p)};")
                        j += 1
                    else:
                        for k in range(pn):
                            out.append(f"{v(r)}.v[{j}] = {v(p)}.v[{k}];")
                            j += 1
                assert j == t.n
                return " ".join(out)
            if t.kind == "struct":
                return " ".join(f"{v(r)}.m{j} = {v(p)};" for j, p in enumerate(parts))
            if t.kind == "array":
                return " ".join(f"{v(r)}.a[{j}] = {v(p)};" for j, p in enumerate(parts))
            raise NotImplementedError(t.kind)
        if op in (81, 82):
            if op == 81:
                rt, r, comp, idx = a[0], a[1], a[2], a[3:]
            else:
                rt, r, obj, comp, idx = a[0], a[1], a[2], a[3], a[4:]
            tid = self.id_type[comp]
            path = ""
            for k in idx:
                t = m.types[tid]
                if t.kind == "struct":
                    path += f".m{k}"
                    tid = t.members[k]
                elif t.kind == "array":
                    path += f".a[{k}]"
                    tid = t.elem
                elif t.kind == "vector":
                    path += f".v[{k}]"
                    tid = t.elem
                else:
                    raise NotImplementedError(t.kind)
            if op == 81:
                return f"{v(r)} = {v(comp)}{path};"
            return f"{v(r)} = {v(comp)}; {v(r)}{path} = {v(obj)};"
        if op == 79:
            rt, r, v1, v2, comps = a[0], a[1], a[2], a[3], a[4:]
            n1 = self.ncomp(self.id_type[v1])
            out = []
            for j, c in enumerate(comps):
                if c == 0xFFFFFFFF:
                    out.append(f"{v(r)}.v[{j}] = 0;")
                elif c < n1:
                    out.append(f"{v(r)}.v[{j}] = {v(v1)}.v[{c}];")
                else:
                    out.append(f"{v(r)}.v[{j}] = {v(v2)}.v[{c - n1}];")
            return " ".join(out)
        if op == 86:
            return f"{v(a[1])}.image = {v(a[2])}; {v(a[1])}.sampler = {v(a[3])};"
        if op in (87, 88):
            rt, r, si, coord = a[0], a[1], a[2], a[3]
            n = self.ncomp(self.id_type[coord])
            lod = "0, 0.0f"
            if op == 88:
                assert a[4] == 2, "only the Lod image operand is supported"
                lod = f"1, {v(a[5])}"
            elif len(a) > 4:
                raise NotImplementedError("image operands on an implicit-lod sample")
            return f"ctx->sample(ctx, {v(si)}.image, {v(si)}.sampler, {v(coord)}.v, {n}, {lod}, {v(r)}.v);"
        if op in (207, 208):
            rt, r, x = a[0], a[1], a[2]
            loc = -1
            src = self.load_src.get(x)
            if src is not None and self.var_meta.get(src, ("",))[0] == "loc":
                loc = self.var_meta[src][1]
            n = self.ncomp(rt)
            xp = f"{v(x)}.v" if n else f"&{v(x)}"
            rp = f"{v(r)}.v" if n else f"&{v(r)}"
            return f"ctx->dpd(ctx, {1 if op == 208 else 0}, {loc}, {max(n, 1)}, {xp}, {rp});"
        if op == 12:
            rt, r, inst, ops = a[0], a[1], a[3], a[4:]
            assert m.ext_sets[a[2]] == "GLSL.std.450"
            fmt = GLSL[inst]
            return self.elementwise(r, rt, fmt, *ops)
        # ---- SPV_KHR_ray_query: the query object keeps what Initialize was given; traversal belongs to the environment
        if op == 4447:   # OpConvertUToAccelerationStructureKHR
            return f"{v(a[1])} = (uint64_t){v(a[2])};"
        if op == 4473:   # OpRayQueryInitializeKHR: query, accel, flags, cull mask, origin, tmin, direction, tmax
            q, acc, flags, mask, org, tmin, dr, tmax = a
            return (f"{v(q)}->accel = {v(acc)}; {v(q)}->flags = {v(flags)}; {v(q)}->cull_mask = {v(mask)}; "
                    f"memcpy({v(q)}->origin, {v(org)}.v, 12); {v(q)}->t_min = {v(tmin)}; memcpy({v(q)}->direction, {v(dr)}.v, 12); "
                    f"{v(q)}->t_max = {v(tmax)}; {v(q)}->initialised = 1u;")
        if op == 4477:   # OpRayQueryProceedKHR
            return f"{v(a[1])} = ctx->rq_proceed(ctx, {v(a[2])});"
        if op == 4476:   # OpRayQueryConfirmIntersectionKHR (only reached while rq_proceed reports candidates)
            return ";"
        if op == 4479:   # OpRayQueryGetIntersectionTypeKHR
            assert m.consts[a[3]][1] == 1, "only the committed intersection is queried by these modules"
            return f"{v(a[1])} = ctx->rq_committed_type(ctx, {v(a[2])});"
        if op == 6019:   # OpRayQueryGetIntersectionInstanceCustomIndexKHR (candidate): read and dropped by the module
            return f"{v(a[1])} = 0u;"
        if op == 232:
            rt, r, p = a[0], a[1], a[2]  # OpAtomicIIncrement; invocations run one at a time
            ld, st = self.loader(rt), self.loader(rt, store=True)
            return f"{v(r)} = {ld}({v(p)}); {st}({v(p)}, {v(r)} + 1u);"
        if op == 124:
            return f"memcpy(&{v(a[1])}, &{v(a[2])}, sizeof {v(a[1])});"
        if op == 169:
            rt, r, c, x, y = a
            if self.ncomp(self.id_type[c]):
                return self.elementwise(r, rt, "({0} ? {1} : {2})", c, x, y)
            return f"{v(r)} = {v(c)} ? {v(x)} : {v(y)};"
        if op == 142:
            rt, r, vec, s = a
            return self.elementwise(r, rt, "{0} * {1}", vec, s)
        if op in UNARY:
            return self.elementwise(a[1], a[0], UNARY[op].replace("RT", self.ctype(self.scalar_of(a[0]))), a[2])
        if op in BINARY:
            return self.elementwise(a[1], a[0], BINARY[op].replace("RT", self.ctype(self.scalar_of(a[0]))), a[2], a[3])
        raise NotImplementedError(f"opcode {op}")

    def scalar_of(self, tid):
        t = self.m.types[tid]
        return t.elem if t.kind == "vector" else tid


RESULT_OPS = {12, 59, 61, 65, 68, 79, 80, 81, 82, 86, 87, 88, 109, 110, 111, 112, 124, 126, 127, 128, 129, 130, 131, 132, 133,
              134, 136, 137, 142, 164, 165, 166, 167, 168, 169, 170, 171, 172, 174, 176, 178, 180, 182, 183, 184, 185, 186,
              187, 188, 189, 190, 191, 194, 196, 197, 198, 199, 207, 208, 232, 4447, 4477, 4479, 6019}
UNARY = {109: "spv_f2u({0})", 110: "spv_f2s({0})", 111: "(float)(int32_t){0}", 112: "(float)(uint32_t){0}",
         126: "(RT)(0u - (uint32_t){0})", 127: "-{0}", 168: "!{0}"}
BINARY = {128: "(RT)((uint32_t){0} + (uint32_t){1})", 130: "(RT)((uint32_t){0} - (uint32_t){1})",
          132: "(RT)((uint32_t){0} * (uint32_t){1})", 134: "(RT)((uint32_t){0} / (uint32_t){1})",
          137: "(RT)((uint32_t){0} % (uint32_t){1})",
          129: "{0} + {1}", 131: "{0} - {1}", 133: "{0} * {1}", 136: "{0} / {1}",
          164: "({0} == {1})", 165: "({0} != {1})", 166: "({0} || {1})", 167: "({0} && {1})",
          170: "((uint32_t){0} == (uint32_t){1})", 171: "((uint32_t){0} != (uint32_t){1})",
          172: "((uint32_t){0} > (uint32_t){1})", 174: "((uint32_t){0} >= (uint32_t){1})",
          176: "((uint32_t){0} < (uint32_t){1})", 178: "((uint32_t){0} <= (uint32_t){1})",
          180: "({0} == {1})", 182: "(({0} < {1}) || ({0} > {1}))", 183: "({0} != {1})", 184: "({0} < {1})",
          185: "(!({0} >= {1}))", 186: "({0} > {1})", 187: "(!({0} <= {1}))", 188: "({0} <= {1})",
          189: "(!({0} > {1}))", 190: "({0} >= {1})", 191: "(!({0} < {1}))",
          194: "(RT)((uint32_t){0} >> ((uint32_t){1} & 31u))", 196: "(RT)((uint32_t){0} << ((uint32_t){1} & 31u))",
          197: "(RT)((uint32_t){0} | (uint32_t){1})", 198: "(RT)((uint32_t){0} ^ (uint32_t){1})",
          199: "(RT)((uint32_t){0} & (uint32_t){1})"}
GLSL = {4: "fabsf({0})", 8: "floorf({0})", 13: "SPV_SIN({0})", 14: "SPV_COS({0})", 26: "powf({0}, {1})", 27: "expf({0})",
        28: "logf({0})", 29: "exp2f({0})", 30: "SPV_LOG2({0})", 31: "sqrtf({0})", 32: "(1.0f / sqrtf({0}))",
        37: "spv_fmin({0}, {1})", 40: "spv_fmax({0}, {1})", 43: "spv_fmin(spv_fmax({0}, {1}), {2})",
        46: "({0} * (1.0f - {2}) + {1} * {2})"}

PRELUDE = r"""/* GENERATED by oracle/spv2c.py from a SPIR-V module shipped by the reference — do not edit, do not commit. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../spv_ctx.h"
/* GLSL.std.450 leaves the precision of these to the implementation; the environment (oracle/build_ref.py) picks the
 * evaluation the C oracle and the CUDA kernels define for them (DESIGN.md "discrete decisions") */
#ifdef SPV_ORACLE_MATH
float orc_log2_spec(float);
#define SPV_LOG2(x) orc_log2_spec(x)
#define SPV_SIN(x) ((float)sin((double)(x)))
#define SPV_COS(x) ((float)cos((double)(x)))
#else
#define SPV_LOG2(x) log2f(x)
#define SPV_SIN(x) sinf(x)
#define SPV_COS(x) cosf(x)
#endif
static inline float spv_bits2f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
/* GLSL.std.450 FMin/FMax are undefined for NaN; Rust's f32::min/max (what the source calls) ignore a NaN operand */
static inline float spv_fmax(float x, float y) { return (x < y || x != x) ? y : x; }
static inline float spv_fmin(float x, float y) { return (y < x || x != x) ? y : x; }
/* OpConvertFToU/S are undefined out of range; Rust's `as` saturates (NaN -> 0) and rust-gpu lowers to these ops */
static inline uint32_t spv_f2u(float f) { return !(f > 0.0f) ? 0u : (f >= 4294967296.0f ? 0xffffffffu : (uint32_t)f); }
static inline int32_t spv_f2s(float f) { return f != f ? 0 : (f <= -2147483648.0f ? INT32_MIN : (f >= 2147483648.0f ? INT32_MAX : (int32_t)f)); }
"""


def translate(spv_path, symbol):
    d = open(spv_path, "rb").read()
    words = struct.unpack(f"<{len(d) // 4}I", d)
    return Emitter(Module(words), symbol).run()


def layouts(spv_path):
    """{struct type id: {"members": [{"offset": o, "type": description}], ...}} plus array strides — the explicit
    layout decorations (OpMemberDecorate Offset / OpDecorate ArrayStride) of a shipped module, for the ABI test."""
    d = open(spv_path, "rb").read()
    m = Module(struct.unpack(f"<{len(d) // 4}I", d))

    def describe(tid):
        t = m.types[tid]
        if t.kind == "int":
            return "i32" if t.signed else "u32"
        if t.kind == "float":
            return "f32"
        if t.kind == "bool":
            return "bool"
        if t.kind == "vector":
            return f"{describe(t.elem)}x{t.n}"
        if t.kind == "array":
            return {"array": describe(t.elem), "len": m.array_len(t), "stride": m.deco.get(tid, {}).get(D_ARRAY_STRIDE, [None])[0]}
        if t.kind == "rtarray":
            return {"rtarray": describe(t.elem), "stride": m.deco.get(tid, {}).get(D_ARRAY_STRIDE, [None])[0]}
        if t.kind == "struct":
            return {"struct": [{"offset": m.mdeco.get((tid, j), {}).get(D_OFFSET, [None])[0], "type": describe(mt)}
                               for j, mt in enumerate(t.members)]}
        return t.kind

    out = []
    for vid, ptid, sc in m.globals:
        if sc in MEMORY_CLASSES:
            d = m.deco.get(vid, {})
            out.append({"storage_class": sc, "set": d.get(D_SET, [None])[0], "binding": d.get(D_BINDING, [None])[0],
                        "type": describe(m.types[ptid].pointee)})
    return {"entry_point": m.entry_name, "interface": out}


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    sym = sys.argv[3] if len(sys.argv) > 3 else "spv_entry"
    open(dst, "w").write(translate(src, sym))
