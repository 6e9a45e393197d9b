"""minijs — a small interpreter for the TypeScript subset the reference's APO hot path is written in.

TEST INFRASTRUCTURE ONLY.  Purpose: PIN the oracle.  This container has no JS runtime (node, deno, bun, quickjs: none), so
the reference's own functions — `_computeRewardSignals` (TCS:668-788), `getStats` (TCS:577-628), `_buildReport`,
`_extractMode`, `_analyzePatterns`, `_generateLocalSuggestions` (APO:498-862), `getStats` (APO:1470-1508) — cannot be run by
a stock engine here.  This module executes their UNMODIFIED source text (read from /root/reference at fixture-generation
time, never copied into the repo) directly: a tokenizer, a recursive-descent parser that skips TypeScript type syntax
(annotations, `as T`, generics, return types) and a tree-walking evaluator with JavaScript semantics for what that text uses:

  * numbers are IEEE-754 binary64 (Python float: same arithmetic, never fused); `/` by zero gives +-Infinity / NaN;
  * `undefined` and `null` are distinct; truthiness, `===` / `!==`, `&&` / `||` / `??` value semantics, `?.` short-circuit;
  * objects keep insertion order (Object.values / Object.entries / for-of over Map.values()); arrays with push, filter,
    find, map, reduce, slice, sort (stable), length; strings with substring / length; Number.prototype.toFixed with the
    ECMAScript algorithm (exact decimal expansion, round half up); Math.max / min; template literals.

Anything outside the subset raises `JSUnsupported` — it never guesses.  `oracle/ts_harness/run_reference.mjs` runs the same
extracted text under Node >= 18 for an independent confirmation on any machine that has one.
"""
from __future__ import annotations

import math
import re
from decimal import Decimal, ROUND_HALF_UP


class JSUnsupported(Exception):
    pass


class _Undefined:
    _inst = None

    def __new__(cls):
        if cls._inst is None:
            cls._inst = super().__new__(cls)
        return cls._inst

    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


undefined = _Undefined()


class JSObject(dict):
    """A plain JS object: ordered string-keyed properties."""
    __slots__ = ()


class JSArray(list):
    __slots__ = ()


class JSMap:
    def __init__(self, items=()):
        self.d = dict(items)


class JSFunction:
    def __init__(self, params, body, env, is_expr, this):
        self.params, self.body, self.env, self.is_expr, self.this = params, body, env, is_expr, this


class SyncPromise:
    """A settled promise for code that is run synchronously: `then` calls its callback at once, `await` yields the value.
    Enough for data-flow checks of async service methods (the order of microtasks is not modelled)."""
    def __init__(self, value=None, error=None, rejected=False):
        self.value, self.error, self.rejected = value, error, rejected


class NativeFunction:
    def __init__(self, fn, props=None):
        self.fn = fn
        self.props = props or {}                       # static members (`Number.isNaN`), constructors called through `new`


# ------------------------------------------------------------------------------------------------ tokenizer
PUNCT = ["===", "!==", "...", ">>>", "=>", ">=", "<=", "&&", "||", "??", "?.", "++", "--", "+=", "-=", "*=", "/=", "|=", "&=", "==", "!=",
         "<<", ">>", "{", "}", "(", ")", "[", "]", ";", ",", "<", ">", "+", "-", "*", "/", "%", "!", "?", ":", "=", ".", "|", "&", "^"]
_hex = re.compile(r"0[xX][0-9a-fA-F]+")
_num = re.compile(r"(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?")
_ident = re.compile(r"[A-Za-z_$][A-Za-z0-9_$]*")


def tokenize(src: str):
    toks, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if c in " \t\r\n":
            i += 1
            continue
        if src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
            continue
        if src.startswith("/*", i):
            j = src.find("*/", i + 2)
            if j < 0:
                raise JSUnsupported("unterminated comment")
            i = j + 2
            continue
        m = _hex.match(src, i)
        if m:
            toks.append(("num", int(m.group(0), 16)))
            i = m.end()
            continue
        m = _num.match(src, i)
        if m and (c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit())):
            text = m.group(0)
            toks.append(("num", float(text) if any(ch in text for ch in ".eE") else int(text)))
            i = m.end()
            continue
        m = _ident.match(src, i)
        if m:
            toks.append(("id", m.group(0)))
            i = m.end()
            continue
        if c in "'\"":
            j, out = i + 1, []
            while j < n and src[j] != c:
                if src[j] == "\\":
                    j += 1
                    esc = src[j]
                    out.append({"n": "\n", "t": "\t", "r": "\r", "0": "\0"}.get(esc, esc))
                else:
                    out.append(src[j])
                j += 1
            toks.append(("str", "".join(out)))
            i = j + 1
            continue
        if c == "`":
            parts, j, buf = [], i + 1, []
            while j < n and src[j] != "`":
                if src[j] == "\\":
                    j += 1
                    buf.append({"n": "\n", "t": "\t"}.get(src[j], src[j]))
                    j += 1
                elif src.startswith("${", j):
                    parts.append(("s", "".join(buf)))
                    buf = []
                    depth, k = 1, j + 2
                    while k < n and depth:
                        if src[k] == "{":
                            depth += 1
                        elif src[k] == "}":
                            depth -= 1
                        k += 1
                    parts.append(("e", src[j + 2:k - 1]))
                    j = k
                else:
                    buf.append(src[j])
                    j += 1
            parts.append(("s", "".join(buf)))
            toks.append(("tpl", parts))
            i = j + 1
            continue
        for p in PUNCT:
            if src.startswith(p, i):
                toks.append(("p", p))
                i += len(p)
                break
        else:
            raise JSUnsupported(f"unexpected character {c!r} at {i}")
    toks.append(("eof", None))
    return toks


# ------------------------------------------------------------------------------------------------ parser
class Parser:
    def __init__(self, src: str):
        self.t = tokenize(src)
        self.i = 0

    # -- token helpers
    def peek(self, k=0):
        return self.t[min(self.i + k, len(self.t) - 1)]

    def at(self, val, k=0):
        tk = self.peek(k)
        return tk[0] in ("p", "id") and tk[1] == val

    def eat(self, val):
        if self.at(val):
            self.i += 1
            return True
        return False

    def need(self, val):
        if not self.eat(val):
            raise JSUnsupported(f"expected {val!r}, got {self.peek()!r} (token {self.i})")

    def ident(self):
        tk = self.peek()
        if tk[0] != "id":
            raise JSUnsupported(f"expected identifier, got {tk!r}")
        self.i += 1
        return tk[1]

    # -- TypeScript type syntax is parsed only to be skipped
    def skip_type(self, stop):
        """Consume a type expression; stops (without consuming) at a token of `stop` at nesting depth 0."""
        depth = 0
        while True:
            tk = self.peek()
            if tk[0] == "eof":
                raise JSUnsupported("unterminated type")
            if tk[0] == "p":
                v = tk[1]
                if v in (">>", ">>>"):                  # inside a type these are closing angle brackets, not shifts
                    self.t[self.i:self.i + 1] = [("p", ">")] * len(v)
                    continue
                if depth == 0 and v in stop:
                    return
                if v in "({[<":
                    depth += 1
                elif v in ")}]>":
                    if depth == 0:
                        return                       # closer of an enclosing construct
                    depth -= 1
                elif v == "=>" and depth == 0 and "=>" in stop:
                    return
            self.i += 1

    # -- statements
    def parse_program(self):
        body = []
        while self.peek()[0] != "eof":
            body.append(self.statement())
        return ("block", body)

    def block(self):
        self.need("{")
        body = []
        while not self.at("}"):
            body.append(self.statement())
        self.need("}")
        return ("block", body)

    def statement(self):
        tk = self.peek()
        if tk == ("p", "{"):
            return self.block()
        if tk == ("p", ";"):
            self.i += 1
            return ("empty",)
        if tk[0] == "id":
            kw = tk[1]
            if kw in ("const", "let", "var"):
                d = self.var_decl()
                self.eat(";")
                return d
            if kw == "if":
                self.i += 1
                self.need("(")
                c = self.expression()
                self.need(")")
                a = self.statement()
                b = None
                if self.at("else"):
                    self.i += 1
                    b = self.statement()
                return ("if", c, a, b)
            if kw == "for":
                return self.for_stmt()
            if kw == "while":
                self.i += 1
                self.need("(")
                c = self.expression()
                self.need(")")
                return ("while", c, self.statement())
            if kw == "return":
                self.i += 1
                e = None if (self.at(";") or self.at("}")) else self.expression()
                self.eat(";")
                return ("return", e)
            if kw == "continue":
                self.i += 1
                self.eat(";")
                return ("continue",)
            if kw == "break":
                self.i += 1
                self.eat(";")
                return ("break",)
            if kw == "try":
                self.i += 1
                b = self.block()
                handler, param, fin = None, None, None
                if self.at("catch"):
                    self.i += 1
                    if self.eat("("):
                        param = self.ident()
                        if self.eat(":"):
                            self.skip_type((")",))
                        self.need(")")
                    handler = self.block()
                if self.at("finally"):
                    self.i += 1
                    fin = self.block()
                return ("try", b, param, handler, fin)
            if kw == "throw":
                self.i += 1
                e = self.expression()
                self.eat(";")
                return ("throw", e)
        e = self.expression()
        self.eat(";")
        return ("expr", e)

    def binding(self):
        """identifier | [a, b] | {a, b: c}"""
        if self.eat("["):
            names = []
            while not self.at("]"):
                names.append(self.binding() if not self.at(",") else None)
                if not self.eat(","):
                    break
            self.need("]")
            return ("arrpat", names)
        if self.eat("{"):
            props = []
            while not self.at("}"):
                key = self.ident()
                target = ("name", key)
                if self.eat(":"):
                    target = self.binding()
                props.append((key, target))
                if not self.eat(","):
                    break
            self.need("}")
            return ("objpat", props)
        return ("name", self.ident())

    def var_decl(self):
        kind = self.ident()
        decls = []
        while True:
            pat = self.binding()
            if self.eat(":"):
                self.skip_type(("=", ",", ";", "of", "in"))
            init = None
            if self.eat("="):
                init = self.assignment()
            decls.append((pat, init))
            if not self.eat(","):
                break
        return ("var", kind, decls)

    def for_stmt(self):
        self.need("for")
        self.need("(")
        if self.peek()[0] == "id" and self.peek()[1] in ("const", "let", "var"):
            save = self.i
            self.i += 1
            pat = self.binding()
            if self.eat(":"):
                self.skip_type(("of", "in", "=", ";"))
            if self.at("of"):
                self.i += 1
                it = self.expression()
                self.need(")")
                return ("forof", pat, it, self.statement())
            self.i = save
            init = self.var_decl()
        else:
            init = None if self.at(";") else ("expr", self.expression())
        self.need(";")
        cond = None if self.at(";") else self.expression()
        self.need(";")
        step = None if self.at(")") else self.expression()
        self.need(")")
        return ("for", init, cond, step, self.statement())

    # -- expressions
    def expression(self):
        e = self.assignment()
        while self.at(","):
            self.i += 1
            e = ("seq", e, self.assignment())
        return e

    def is_arrow_ahead(self):
        """At '(' : is this `(params) [: type] =>` ?"""
        depth, k = 0, self.i
        while True:
            tk = self.t[k]
            if tk[0] == "eof":
                return False
            if tk[0] == "p":
                if tk[1] in "([{":
                    depth += 1
                elif tk[1] in ")]}":
                    depth -= 1
                    if depth == 0:
                        nxt = self.t[k + 1]
                        if nxt == ("p", "=>"):
                            return True
                        if nxt == ("p", ":"):
                            # return type annotation: scan to => at depth 0 before a statement boundary
                            d2, j = 0, k + 2
                            while self.t[j][0] != "eof":
                                v = self.t[j]
                                if v[0] == "p":
                                    if v[1] in "([{<":
                                        d2 += 1
                                    elif v[1] in ")]}>":
                                        if d2 == 0:
                                            return False
                                        d2 -= 1
                                    elif v[1] == "=>" and d2 == 0:
                                        return True
                                    elif v[1] in (";", ",", "=") and d2 == 0:
                                        return False
                                j += 1
                        return False
            k += 1

    def arrow(self):
        params = []
        if self.peek()[0] == "id" and self.peek(1) == ("p", "=>"):
            params.append(("name", self.ident()))
        else:
            self.need("(")
            while not self.at(")"):
                p = self.binding()
                self.eat("?")
                if self.eat(":"):
                    self.skip_type((",", ")", "="))
                if self.eat("="):
                    self.assignment()            # default values are not used by the path
                    raise JSUnsupported("default parameter values")
                params.append(p)
                if not self.eat(","):
                    break
            self.need(")")
            if self.eat(":"):
                self.skip_type(("=>",))
        self.need("=>")
        if self.at("{"):
            return ("fn", params, self.block(), False)
        return ("fn", params, self.assignment(), True)

    def assignment(self):
        tk = self.peek()
        if tk[0] == "id" and self.peek(1) == ("p", "=>") and tk[1] not in ("return",):
            return self.arrow()
        if tk == ("p", "(") and self.is_arrow_ahead():
            return self.arrow()
        left = self.ternary()
        for op in ("=", "+=", "-=", "*=", "/=", "|=", "&="):
            if self.at(op) and self.peek()[0] == "p":
                self.i += 1
                right = self.assignment()
                if left[0] not in ("name", "member", "index"):
                    raise JSUnsupported("assignment target")
                return ("assign", op, left, right)
        return left

    def ternary(self):
        c = self.nullish()
        if self.at("?") and self.peek()[0] == "p":
            self.i += 1
            a = self.assignment()
            self.need(":")
            b = self.assignment()
            return ("cond", c, a, b)
        return c

    def nullish(self):
        e = self.logical_or()
        while self.at("??"):
            self.i += 1
            e = ("nullish", e, self.logical_or())
        return e

    def logical_or(self):
        e = self.logical_and()
        while self.at("||"):
            self.i += 1
            e = ("or", e, self.logical_and())
        return e

    def logical_and(self):
        e = self.bit_or()
        while self.at("&&"):
            self.i += 1
            e = ("and", e, self.bit_or())
        return e

    def bit_or(self):
        e = self.bit_xor()
        while self.at("|") and self.peek()[0] == "p":
            self.i += 1
            e = ("bin", "|", e, self.bit_xor())
        return e

    def bit_xor(self):
        e = self.bit_and()
        while self.at("^") and self.peek()[0] == "p":
            self.i += 1
            e = ("bin", "^", e, self.bit_and())
        return e

    def bit_and(self):
        e = self.equality()
        while self.at("&") and self.peek()[0] == "p":
            self.i += 1
            e = ("bin", "&", e, self.equality())
        return e

    def equality(self):
        e = self.relational()
        while self.peek()[0] == "p" and self.peek()[1] in ("===", "!==", "==", "!="):
            op = self.peek()[1]
            self.i += 1
            e = ("bin", op, e, self.relational())
        return e

    def shift(self):
        e = self.additive()
        while self.peek()[0] == "p" and self.peek()[1] in ("<<", ">>", ">>>"):
            op = self.peek()[1]
            self.i += 1
            e = ("bin", op, e, self.additive())
        return e

    def relational(self):
        e = self.shift()
        while True:
            tk = self.peek()
            if tk[0] == "p" and tk[1] in ("<", ">", "<=", ">="):
                self.i += 1
                e = ("bin", tk[1], e, self.shift())
            elif tk == ("id", "as"):
                self.i += 1
                self.skip_type((",", ")", "]", "}", ";", "?", ":", "||", "&&", "??", "===", "!==", "=", "=>", "+", "-"))
            elif tk == ("id", "instanceof"):
                self.i += 1
                e = ("instanceof", e, self.shift())
            elif tk == ("id", "in"):
                raise JSUnsupported(tk[1])
            else:
                return e

    def additive(self):
        e = self.multiplicative()
        while self.peek()[0] == "p" and self.peek()[1] in ("+", "-"):
            op = self.peek()[1]
            self.i += 1
            e = ("bin", op, e, self.multiplicative())
        return e

    def multiplicative(self):
        e = self.unary()
        while self.peek()[0] == "p" and self.peek()[1] in ("*", "/", "%"):
            op = self.peek()[1]
            self.i += 1
            e = ("bin", op, e, self.unary())
        return e

    def unary(self):
        tk = self.peek()
        if tk[0] == "p" and tk[1] in ("!", "-", "+"):
            self.i += 1
            return ("un", tk[1], self.unary())
        if tk[0] == "p" and tk[1] in ("++", "--"):
            self.i += 1
            target = self.unary()
            return ("update", tk[1], target, True)
        if tk == ("id", "typeof"):
            self.i += 1
            return ("typeof", self.unary())
        if tk == ("id", "await"):
            self.i += 1
            return ("await", self.unary())
        if tk[0] == "id" and tk[1] in ("delete", "void", "yield"):
            raise JSUnsupported(tk[1])
        e = self.postfix()
        while self.at("!") and self.peek(1)[0] == "p" and self.peek(1)[1] in (".", ")", ",", ";", "]", "["):
            self.i += 1                                  # TypeScript non-null assertion
        return e

    def postfix(self):
        e = self.call_member()
        tk = self.peek()
        if tk[0] == "p" and tk[1] in ("++", "--"):
            self.i += 1
            return ("update", tk[1], e, False)
        return e

    def arguments(self):
        args = []
        self.need("(")
        while not self.at(")"):
            if self.eat("..."):
                args.append(("spread", self.assignment()))
            else:
                args.append(self.assignment())
            if not self.eat(","):
                break
        self.need(")")
        return args

    def call_member(self):
        e = self.primary()
        while True:
            tk = self.peek()
            if tk == ("p", "."):
                self.i += 1
                e = ("member", e, self.ident(), False)
            elif tk == ("p", "?."):
                self.i += 1
                if self.at("("):
                    e = ("call", e, self.arguments(), True)
                elif self.at("["):
                    self.i += 1
                    k = self.expression()
                    self.need("]")
                    e = ("index", e, k, True)
                else:
                    e = ("member", e, self.ident(), True)
            elif tk == ("p", "["):
                self.i += 1
                k = self.expression()
                self.need("]")
                e = ("index", e, k, False)
            elif tk == ("p", "("):
                e = ("call", e, self.arguments(), False)
            elif tk == ("p", "<") and e[0] in ("name", "member"):
                # generic call `f<T>(...)`: only when a matching '>' is directly followed by '('
                depth, k = 0, self.i
                ok = False
                while self.t[k][0] != "eof":
                    v = self.t[k]
                    if v == ("p", "<"):
                        depth += 1
                    elif v == ("p", ">"):
                        depth -= 1
                        if depth == 0:
                            ok = self.t[k + 1] == ("p", "(")
                            break
                    elif v[0] == "p" and v[1] in (";", "{", "}", "=", "&&", "||", ")"):
                        break
                    k += 1
                if not ok:
                    return e
                self.i = k + 1
            else:
                return e

    def primary(self):
        tk = self.peek()
        kind, v = tk
        if kind == "num" or kind == "str":
            self.i += 1
            return ("lit", v)
        if kind == "tpl":
            self.i += 1
            return ("tpl", [(k, s if k == "s" else Parser(s).expression()) for k, s in v])
        if kind == "p":
            if v == "(":
                self.i += 1
                e = self.expression()
                self.need(")")
                return e
            if v == "[":
                self.i += 1
                items = []
                while not self.at("]"):
                    if self.eat("..."):
                        items.append(("spread", self.assignment()))
                    else:
                        items.append(self.assignment())
                    if not self.eat(","):
                        break
                self.need("]")
                return ("array", items)
            if v == "{":
                self.i += 1
                props = []
                while not self.at("}"):
                    if self.eat("..."):
                        props.append(("spread", self.assignment()))
                    else:
                        k = self.peek()
                        if k[0] in ("id", "str", "num"):
                            self.i += 1
                            key = str(k[1])
                        elif k == ("p", "["):
                            raise JSUnsupported("computed property")
                        else:
                            raise JSUnsupported(f"property key {k!r}")
                        if self.eat(":"):
                            props.append(("kv", key, self.assignment()))
                        elif k[0] == "id":
                            props.append(("kv", key, ("name", key)))       # shorthand
                        else:
                            raise JSUnsupported("property without value")
                    if not self.eat(","):
                        break
                self.need("}")
                return ("object", props)
        if kind == "id":
            if v in ("true", "false"):
                self.i += 1
                return ("lit", v == "true")
            if v == "null":
                self.i += 1
                return ("lit", None)
            if v == "undefined":
                self.i += 1
                return ("lit", undefined)
            if v == "this":
                self.i += 1
                return ("this",)
            if v == "new":
                self.i += 1
                callee = ("name", self.ident())
                while self.eat("."):
                    callee = ("member", callee, self.ident(), False)
                if self.at("<"):
                    self.i += 1
                    self.skip_type(())
                    self.need(">")
                args = self.arguments() if self.at("(") else []
                return ("new", callee, args)
            if v == "function":
                raise JSUnsupported("function expression")
            self.i += 1
            return ("name", v)
        raise JSUnsupported(f"unexpected token {tk!r}")


def parse_method(src: str):
    """`[private] name(params): T { body }` -> (name, [param patterns], body block).  Types are skipped."""
    p = Parser(src)
    while p.peek()[0] == "id" and p.peek()[1] in ("private", "public", "protected", "async", "static", "readonly"):
        p.i += 1                                        # `async`: the body runs synchronously, `await` unwraps a SyncPromise
    name = p.ident()
    p.need("(")
    params = []
    while not p.at(")"):
        b = p.binding()
        p.eat("?")
        if p.eat(":"):
            p.skip_type((",", ")", "="))
        if p.eat("="):
            b = ("defpat", b, p.assignment())           # parameter default: evaluated in the callee's scope when the argument is undefined
        params.append(b)
        if not p.eat(","):
            break
    p.need(")")
    if p.eat(":"):
        p.skip_type(("{",))
    body = p.block()
    if p.peek()[0] != "eof":
        raise JSUnsupported("trailing tokens after method body")
    return name, params, body


# ------------------------------------------------------------------------------------------------ evaluator
class _Return(Exception):
    def __init__(self, v):
        self.v = v


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class JSThrow(Exception):
    def __init__(self, v):
        self.v = v


class Env:
    def __init__(self, parent=None):
        self.vars, self.parent = {}, parent

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e
            e = e.parent
        return None


def truthy(v):
    if v is None or v is undefined or v is False:
        return False
    if v is True:
        return True
    if isinstance(v, (int, float)):
        return not (v == 0 or (isinstance(v, float) and math.isnan(v)))
    if isinstance(v, str):
        return len(v) > 0
    return True


def is_num(v):
    return isinstance(v, (int, float)) and not isinstance(v, bool)


def to_number(v):
    if is_num(v):
        return v
    if v is True:
        return 1
    if v is False or v is None:
        return 0
    if v is undefined:
        return math.nan
    if isinstance(v, str):
        try:
            return float(v) if v.strip() else 0
        except ValueError:
            return math.nan
    raise JSUnsupported(f"ToNumber of {type(v).__name__}")


def num_to_string(x):
    """Number::toString for the values this path prints (integers and short decimals)."""
    if isinstance(x, bool):
        return "true" if x else "false"
    if isinstance(x, int):
        return str(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == int(x) and abs(x) < 1e21:
        return str(int(x))
    r = repr(x)                                    # shortest round-trip digits, as ECMAScript specifies
    if "e" in r or "E" in r:
        raise JSUnsupported("exponent formatting in Number::toString")
    return r


def to_string(v):
    if isinstance(v, str):
        return v
    if v is None:
        return "null"
    if v is undefined:
        return "undefined"
    if isinstance(v, bool) or is_num(v):
        return num_to_string(v)
    raise JSUnsupported(f"ToString of {type(v).__name__}")


def to_fixed(x, digits):
    """Number.prototype.toFixed (ECMAScript 21.1.3.3): n such that n / 10^f - x is as close to zero as possible, the larger n
    on a tie — evaluated on the exact binary value of x."""
    x = float(to_number(x))
    if math.isnan(x):
        return "NaN"
    if abs(x) >= 1e21:
        return num_to_string(x)
    q = Decimal(1).scaleb(-int(digits))
    d = Decimal(abs(x)).quantize(q, rounding=ROUND_HALF_UP)
    s = format(d, "f")
    return ("-" if x < 0 else "") + s          # the sign is that of x (ECMAScript step 5): (-0.0004).toFixed(3) is "-0.000"


def to_int32(v):
    x = to_number(v)
    if isinstance(x, bool):
        x = int(x)
    if isinstance(x, float):
        if math.isnan(x) or math.isinf(x):
            return 0
        x = int(x)                                        # truncation toward zero
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def strict_equal(a, b):
    if is_num(a) and is_num(b):
        return float(a) == float(b)
    if type(a) is not type(b) and not (isinstance(a, str) and isinstance(b, str)):
        return False
    if isinstance(a, (str, bool)) or a is None or a is undefined:
        return a == b if not (a is None or a is undefined) else a is b
    return a is b


def js_div(a, b):
    a, b = float(a), float(b)
    if b == 0.0:
        if a == 0.0 or math.isnan(a):
            return math.nan
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf
    return a / b


class Interp:
    def __init__(self, globals_=None):
        self.g = Env()
        self.g.vars.update({
            "Infinity": math.inf, "NaN": math.nan,
            "Math": JSObject(max=NativeFunction(self._max), min=NativeFunction(self._min),
                             abs=NativeFunction(lambda t, x: abs(to_number(x))), floor=NativeFunction(lambda t, x: float(math.floor(to_number(x)))),
                             round=NativeFunction(lambda t, x: float(math.floor(to_number(x) + 0.5)))),
            "Object": JSObject(values=NativeFunction(lambda t, o: JSArray(o.values())),
                               keys=NativeFunction(lambda t, o: JSArray(o.keys())),
                               entries=NativeFunction(lambda t, o: JSArray(JSArray([k, v]) for k, v in o.items()))),
            "Array": JSObject(**{"from": NativeFunction(lambda t, it: JSArray(self.iterate(it))),
                                 "isArray": NativeFunction(lambda t, v: isinstance(v, JSArray))}),
            "Number": NativeFunction(lambda t, v=0: to_number(v),
                                     {"isNaN": NativeFunction(lambda t, v=undefined: isinstance(v, float) and math.isnan(v)),
                                      "isFinite": NativeFunction(lambda t, v=undefined: is_num(v) and math.isfinite(v))}),
            "String": NativeFunction(lambda t, v="": to_string(v)),
            "Boolean": NativeFunction(lambda t, v=False: truthy(v)),
        })
        if globals_:
            self.g.vars.update(globals_)

    @staticmethod
    def _max(this, *a):
        r = -math.inf
        for v in a:
            v = to_number(v)
            if isinstance(v, float) and math.isnan(v):
                return math.nan
            if v > r or (v == 0 and r == 0 and math.copysign(1.0, float(r)) < 0):
                r = v
        return r

    @staticmethod
    def _min(this, *a):
        r = math.inf
        for v in a:
            v = to_number(v)
            if isinstance(v, float) and math.isnan(v):
                return math.nan
            if v < r:
                r = v
        return r

    # -- calling
    def make_method(self, src: str, this):
        name, params, body = parse_method(src)
        return name, JSFunction(params, body, self.g, False, this)

    def call(self, fn, this, args):
        if isinstance(fn, NativeFunction):
            return fn.fn(this, *args)
        if not isinstance(fn, JSFunction):
            raise JSThrow(f"TypeError: {fn!r} is not a function")
        env = Env(fn.env)
        for k, p in enumerate(fn.params):
            self.bind(env, p, args[k] if k < len(args) else undefined, "let")
        this_v = fn.this                                   # arrows and the extracted methods carry their `this`
        if fn.is_expr:
            return self.ev_top(fn.body, env, this_v)
        try:
            self.exec(fn.body, env, this_v)
        except _Return as r:
            return r.v
        return undefined

    def bind(self, env, pat, value, kind):
        if pat[0] == "name":
            env.vars[pat[1]] = value
        elif pat[0] == "arrpat":
            seq = list(self.iterate(value))
            for k, sub in enumerate(pat[1]):
                if sub is not None:
                    self.bind(env, sub, seq[k] if k < len(seq) else undefined, kind)
        elif pat[0] == "objpat":
            for key, target in pat[1]:
                self.bind(env, target, self.get(value, key), kind)
        elif pat[0] == "defpat":
            self.bind(env, pat[1], self.ev_top(pat[2], env, undefined) if value is undefined else value, kind)
        else:
            raise JSUnsupported(pat[0])

    def iterate(self, v):
        if isinstance(v, JSArray):
            i = 0
            while i < len(v):                              # live, like a JS array iterator
                yield v[i]
                i += 1
        elif isinstance(v, str):
            yield from v
        elif isinstance(v, (list, tuple)):
            yield from v
        elif hasattr(v, "__iter__") and not isinstance(v, (dict, JSMap)):
            yield from v
        else:
            raise JSThrow(f"TypeError: {type(v).__name__} is not iterable")

    # -- property access
    def get(self, obj, key):
        if obj is None or obj is undefined:
            raise JSThrow(f"TypeError: cannot read properties of {obj!r} (reading {key!r})")
        if isinstance(obj, JSArray):
            if isinstance(key, (int, float)) and not isinstance(key, bool):
                k = int(key)
                return obj[k] if 0 <= k < len(obj) and k == key else undefined
            if key == "length":
                return len(obj)
            return self.array_method(obj, key)
        if isinstance(obj, JSObject):
            if key in obj:
                return obj[key]
            return undefined
        if isinstance(obj, str):
            if key == "length":
                return len(obj)
            if isinstance(key, (int, float)):
                return obj[int(key)] if 0 <= int(key) < len(obj) else undefined
            return self.string_method(obj, key)
        if is_num(obj):
            if key == "toFixed":
                return NativeFunction(lambda this, d=0: to_fixed(obj, to_number(d)))
            if key == "toString":
                return NativeFunction(lambda this: num_to_string(obj))
            raise JSUnsupported(f"Number.prototype.{key}")
        if isinstance(obj, SyncPromise):
            def then(this, f=undefined, g=undefined):
                try:
                    if obj.rejected:
                        return obj if g is undefined else SyncPromise(self.call(g, undefined, [obj.error]))
                    return obj if f is undefined else SyncPromise(self.call(f, undefined, [obj.value]))
                except JSThrow as e:
                    return SyncPromise(error=e.v, rejected=True)
            if key == "then":
                return NativeFunction(then)
            if key == "catch":
                return NativeFunction(lambda this, g=undefined: then(this, undefined, g))
            raise JSUnsupported(f"Promise.prototype.{key}")
        if isinstance(obj, NativeFunction):
            if key in obj.props:
                return obj.props[key]
            raise JSUnsupported(f"property {key!r} of a native function")
        if isinstance(obj, JSMap):
            if key == "size":
                return len(obj.d)
            if key == "values":
                return NativeFunction(lambda this: JSArray(obj.d.values()))
            if key == "keys":
                return NativeFunction(lambda this: JSArray(obj.d.keys()))
            if key == "get":
                return NativeFunction(lambda this, k: obj.d.get(k, undefined))
            if key == "has":
                return NativeFunction(lambda this, k: k in obj.d)
            raise JSUnsupported(f"Map.prototype.{key}")
        raise JSUnsupported(f"property {key!r} of {type(obj).__name__}")

    def array_method(self, arr, name):
        call = self.call
        if name == "push":
            def push(this, *a):
                arr.extend(a)
                return len(arr)
            return NativeFunction(push)
        if name == "filter":
            return NativeFunction(lambda this, f: JSArray(v for i, v in enumerate(list(arr)) if truthy(call(f, undefined, [v, i, arr]))))
        if name == "map":
            return NativeFunction(lambda this, f: JSArray(call(f, undefined, [v, i, arr]) for i, v in enumerate(list(arr))))
        if name == "forEach":
            def each(this, f):
                for i, v in enumerate(list(arr)):
                    call(f, undefined, [v, i, arr])
                return undefined
            return NativeFunction(each)
        if name == "find":
            def find(this, f):
                for i, v in enumerate(list(arr)):
                    if truthy(call(f, undefined, [v, i, arr])):
                        return v
                return undefined
            return NativeFunction(find)
        if name == "some":
            return NativeFunction(lambda this, f: any(truthy(call(f, undefined, [v, i, arr])) for i, v in enumerate(list(arr))))
        if name == "every":
            return NativeFunction(lambda this, f: all(truthy(call(f, undefined, [v, i, arr])) for i, v in enumerate(list(arr))))
        if name == "reduce":
            def reduce(this, f, *init):
                it = list(arr)
                if init:
                    acc, start = init[0], 0
                else:
                    if not it:
                        raise JSThrow("TypeError: reduce of empty array with no initial value")
                    acc, start = it[0], 1
                for i in range(start, len(it)):
                    acc = call(f, undefined, [acc, it[i], i, arr])
                return acc
            return NativeFunction(reduce)
        if name == "slice":
            def slc(this, a=0, b=undefined):
                n = len(arr)
                a = int(to_number(a))
                b = n if b is undefined else int(to_number(b))
                a = max(n + a, 0) if a < 0 else min(a, n)
                b = max(n + b, 0) if b < 0 else min(b, n)
                return JSArray(arr[a:b])
            return NativeFunction(slc)
        if name == "sort":
            def sort(this, f=undefined):
                import functools
                if f is undefined:
                    raise JSUnsupported("sort without comparator")
                def cmp(a, b):
                    r = to_number(call(f, undefined, [a, b]))
                    return -1 if r < 0 else (1 if r > 0 else 0)
                arr.sort(key=functools.cmp_to_key(cmp))    # Python's sort is stable, as Array.prototype.sort is (ES2019)
                return arr
            return NativeFunction(sort)
        if name == "join":
            return NativeFunction(lambda this, sep=",": to_string(sep).join("" if (v is None or v is undefined) else to_string(v) for v in arr))
        if name == "includes":
            return NativeFunction(lambda this, x: any(strict_equal(v, x) for v in arr))
        if name == "indexOf":
            def index_of(this, x):
                for i, v in enumerate(arr):
                    if strict_equal(v, x):
                        return i
                return -1
            return NativeFunction(index_of)
        raise JSUnsupported(f"Array.prototype.{name}")

    def string_method(self, s, name):
        if name == "substring":
            def substring(this, a=0, b=undefined):
                n = len(s)
                a = min(max(int(to_number(a)), 0), n)
                b = n if b is undefined else min(max(int(to_number(b)), 0), n)
                if a > b:
                    a, b = b, a
                return s[a:b]
            return NativeFunction(substring)
        if name == "slice":
            def slc(this, a=0, b=undefined):
                n = len(s)
                a = int(to_number(a))
                b = n if b is undefined else int(to_number(b))
                a = max(n + a, 0) if a < 0 else min(a, n)
                b = max(n + b, 0) if b < 0 else min(b, n)
                return s[a:b]
            return NativeFunction(slc)
        if name == "trim":
            return NativeFunction(lambda this: s.strip())
        if name == "startsWith":
            return NativeFunction(lambda this, p: s.startswith(p))
        if name == "toLowerCase":
            return NativeFunction(lambda this: s.lower())
        raise JSUnsupported(f"String.prototype.{name}")

    # -- statements
    def exec(self, node, env, this):
        k = node[0]
        if k == "block":
            inner = Env(env)
            for st in node[1]:
                self.exec(st, inner, this)
        elif k == "var":
            for pat, init in node[2]:
                self.bind(env, pat, undefined if init is None else self.ev_top(init, env, this), node[1])
        elif k == "expr":
            self.ev_top(node[1], env, this)
        elif k == "if":
            if truthy(self.ev_top(node[1], env, this)):
                self.exec(node[2], env, this)
            elif node[3] is not None:
                self.exec(node[3], env, this)
        elif k == "forof":
            for v in self.iterate(self.ev_top(node[2], env, this)):
                inner = Env(env)
                self.bind(inner, node[1], v, "const")
                try:
                    self.exec(node[3], inner, this)
                except _Continue:
                    continue
                except _Break:
                    break
        elif k == "for":
            inner = Env(env)
            if node[1] is not None:
                self.exec(node[1], inner, this)
            while node[2] is None or truthy(self.ev_top(node[2], inner, this)):
                try:
                    self.exec(node[4], inner, this)
                except _Continue:
                    pass
                except _Break:
                    break
                if node[3] is not None:
                    self.ev_top(node[3], inner, this)
        elif k == "while":
            while truthy(self.ev_top(node[1], env, this)):
                try:
                    self.exec(node[2], env, this)
                except _Continue:
                    continue
                except _Break:
                    break
        elif k == "return":
            raise _Return(undefined if node[1] is None else self.ev_top(node[1], env, this))
        elif k == "continue":
            raise _Continue()
        elif k == "break":
            raise _Break()
        elif k == "try":
            try:
                try:
                    self.exec(node[1], env, this)
                except JSThrow as ex:
                    if node[3] is None:
                        raise
                    inner = Env(env)
                    if node[2]:
                        inner.vars[node[2]] = ex.v
                    self.exec(node[3], inner, this)
            finally:
                if node[4] is not None:
                    self.exec(node[4], env, this)
        elif k == "throw":
            raise JSThrow(self.ev_top(node[1], env, this))
        elif k == "empty":
            pass
        else:
            raise JSUnsupported(f"statement {k}")

    # -- expressions
    def assign_to(self, target, value, env, this):
        if target[0] == "name":
            scope = env.lookup(target[1])
            if scope is None:
                raise JSThrow(f"ReferenceError: {target[1]} is not defined")
            scope.vars[target[1]] = value
        elif target[0] == "member":
            obj = self.ev_top(target[1], env, this)
            if not isinstance(obj, JSObject):
                raise JSUnsupported(f"assignment to a property of {type(obj).__name__}")
            obj[target[2]] = value
        elif target[0] == "index":
            obj = self.ev_top(target[1], env, this)
            key = self.ev_top(target[2], env, this)
            if isinstance(obj, JSObject):
                obj[to_string(key)] = value
            elif isinstance(obj, JSArray) and is_num(key):
                k = int(key)
                while len(obj) <= k:
                    obj.append(undefined)
                obj[k] = value
            else:
                raise JSUnsupported("indexed assignment")
        else:
            raise JSUnsupported("assignment target")

    def binary(self, op, a, b):
        if op == "+":
            if isinstance(a, str) or isinstance(b, str):
                return to_string(a) + to_string(b)
            a, b = to_number(a), to_number(b)
            return a + b
        if op in ("-", "*", "/", "%"):
            a, b = to_number(a), to_number(b)
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            if op == "/":
                return js_div(a, b)
            if b == 0:
                return math.nan
            return math.fmod(a, b)
        if op in ("<", ">", "<=", ">="):
            if isinstance(a, str) and isinstance(b, str):
                pass
            else:
                a, b = to_number(a), to_number(b)
                if (isinstance(a, float) and math.isnan(a)) or (isinstance(b, float) and math.isnan(b)):
                    return False
            return {"<": a < b, ">": a > b, "<=": a <= b, ">=": a >= b}[op]
        if op in ("|", "&", "^", "<<", ">>", ">>>"):
            x, y = to_int32(a), to_int32(b)
            if op == "|":
                r = x | y
            elif op == "&":
                r = x & y
            elif op == "^":
                r = x ^ y
            elif op == "<<":
                r = x << (y & 31)
            elif op == ">>":
                r = x >> (y & 31)
            else:
                return (x & 0xFFFFFFFF) >> (y & 31)
            r &= 0xFFFFFFFF
            return r - (1 << 32) if r & 0x80000000 else r
        if op == "===":
            return strict_equal(a, b)
        if op == "!==":
            return not strict_equal(a, b)
        if op in ("==", "!="):
            nullish_a, nullish_b = a is None or a is undefined, b is None or b is undefined
            if nullish_a or nullish_b:
                r = nullish_a and nullish_b
            elif type(a) is type(b) or (is_num(a) and is_num(b)):
                r = strict_equal(a, b)
            else:
                raise JSUnsupported("loose equality between different types")
            return r if op == "==" else not r
        raise JSUnsupported(op)

    def ev(self, node, env, this):
        k = node[0]
        if k == "lit":
            return node[1]
        if k == "name":
            scope = env.lookup(node[1])
            if scope is None:
                raise JSThrow(f"ReferenceError: {node[1]} is not defined")
            return scope.vars[node[1]]
        if k == "this":
            return this
        if k == "member":
            obj = self.ev(node[1], env, this)
            if node[3] and (obj is None or obj is undefined):
                raise _ShortCircuit()
            return self.get(obj, node[2])
        if k == "index":
            obj = self.ev(node[1], env, this)
            if node[3] and (obj is None or obj is undefined):
                raise _ShortCircuit()
            key = self.ev_top(node[2], env, this)
            return self.get(obj, key if is_num(key) else to_string(key))
        if k == "call":
            callee = node[1]
            if callee[0] in ("member", "index"):
                obj = self.ev(callee[1], env, this)
                if callee[3] and (obj is None or obj is undefined):
                    raise _ShortCircuit()
                key = callee[2] if callee[0] == "member" else self.ev_top(callee[2], env, this)
                fn, this_v = self.get(obj, key), obj
            else:
                fn, this_v = self.ev(callee, env, this), undefined
            if node[3] and (fn is None or fn is undefined):
                raise _ShortCircuit()
            args = []
            for a in node[2]:
                if a[0] == "spread":
                    args.extend(self.iterate(self.ev_top(a[1], env, this)))
                else:
                    args.append(self.ev_top(a, env, this))
            return self.call(fn, this_v, args)
        if k == "fn":
            return JSFunction(node[1], node[2], env, node[3], this)
        if k == "object":
            o = JSObject()
            for p in node[1]:
                if p[0] == "spread":
                    src = self.ev_top(p[1], env, this)
                    if isinstance(src, JSObject):
                        o.update(src)
                    elif not (src is None or src is undefined):
                        raise JSUnsupported("object spread of a non-object")
                else:
                    o[p[1]] = self.ev_top(p[2], env, this)
            return o
        if k == "array":
            a = JSArray()
            for it in node[1]:
                if it[0] == "spread":
                    a.extend(self.iterate(self.ev_top(it[1], env, this)))
                else:
                    a.append(self.ev_top(it, env, this))
            return a
        if k == "tpl":
            return "".join(s if kind == "s" else to_string(self.ev_top(s, env, this)) for kind, s in node[1])
        if k == "cond":
            return self.ev_top(node[2] if truthy(self.ev_top(node[1], env, this)) else node[3], env, this)
        if k == "and":
            a = self.ev_top(node[1], env, this)
            return self.ev_top(node[2], env, this) if truthy(a) else a
        if k == "or":
            a = self.ev_top(node[1], env, this)
            return a if truthy(a) else self.ev_top(node[2], env, this)
        if k == "nullish":
            a = self.ev_top(node[1], env, this)
            return self.ev_top(node[2], env, this) if (a is None or a is undefined) else a
        if k == "bin":
            a = self.ev_top(node[2], env, this)
            b = self.ev_top(node[3], env, this)
            return self.binary(node[1], a, b)
        if k == "un":
            v = self.ev_top(node[2], env, this)
            if node[1] == "!":
                return not truthy(v)
            v = to_number(v)
            return -v if node[1] == "-" else v
        if k == "await":
            v = self.ev_top(node[1], env, this)
            if isinstance(v, SyncPromise):
                if v.rejected:
                    raise JSThrow(v.error)
                return v.value
            return v
        if k == "instanceof":
            v = self.ev_top(node[1], env, this)
            ctor = self.ev_top(node[2], env, this)
            return isinstance(v, JSObject) and v.get("__ctor__", undefined) is ctor
        if k == "typeof":
            v = self.ev_top(node[1], env, this)
            if v is undefined:
                return "undefined"
            if v is None or isinstance(v, (JSObject, JSArray, JSMap)):
                return "object"
            if isinstance(v, bool):
                return "boolean"
            if is_num(v):
                return "number"
            if isinstance(v, str):
                return "string"
            return "function"
        if k == "update":
            old = to_number(self.ev_top(node[2], env, this))
            new = old + 1 if node[1] == "++" else old - 1
            self.assign_to(node[2], new, env, this)
            return new if node[3] else old
        if k == "assign":
            op, target = node[1], node[2]
            if op == "=":
                v = self.ev_top(node[3], env, this)
            else:
                cur = self.ev_top(target, env, this)
                v = self.binary(op[0], cur, self.ev_top(node[3], env, this))
            self.assign_to(target, v, env, this)
            return v
        if k == "seq":
            self.ev_top(node[1], env, this)
            return self.ev_top(node[2], env, this)
        if k == "new":
            ctor = node[1]
            if ctor == ("name", "Map"):
                args = [self.ev_top(a, env, this) for a in node[2]]
                return JSMap((e[0], e[1]) for e in (args[0] if args else []))
            if ctor[0] == "name":
                scope = env.lookup(ctor[1])
                fn = scope.vars[ctor[1]] if scope is not None else None
                if isinstance(fn, NativeFunction):         # host-provided constructor (DataView in the codec tests)
                    return fn.fn(undefined, *[self.ev_top(a, env, this) for a in node[2]])
            raise JSUnsupported(f"new {ctor}")
        raise JSUnsupported(f"expression {k}")

    def ev_top(self, node, env, this):
        """Evaluates one complete optional chain: `a?.b.c` yields undefined as a whole when `a` is nullish."""
        try:
            return self.ev(node, env, this)
        except _ShortCircuit:
            return undefined


class _ShortCircuit(Exception):
    pass


# ------------------------------------------------------------------------------------------------ Python <-> JS values
def to_js(v):
    if isinstance(v, dict):
        return JSObject((k, to_js(x)) for k, x in v.items())
    if isinstance(v, (list, tuple)):
        return JSArray(to_js(x) for x in v)
    return v


def from_js(v):
    if isinstance(v, JSObject):
        return {k: from_js(x) for k, x in v.items()}
    if isinstance(v, JSArray):
        return [from_js(x) for x in v]
    if v is undefined:
        return None
    return v
