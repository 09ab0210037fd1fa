#!/usr/bin/env python
"""Generates etl_b200/csrc/json_tables.cuh: the byte-class and transition tables of the table-driven
JSON acceptor used by k_walk (serde_json acceptance, what `serde_json::from_str::<Value>` accepts —
conversions/text.rs:104-107 in the reference; grammar restated in oracle/oracle_cells.c).

The acceptor is written so that every lane of a warp executes the same instructions per byte:
    cls = CLS[byte];  e = TRANS[state * NCLS + cls];  state = e & 31;  action = e >> 5
plus branch-free handling of the 7 actions (container stack, comma, string entry).  Escapes inside
strings (\\uXXXX, surrogate pairs) are rare and handled by a small divergent block on the device.

`simulate()` is the same algorithm in Python; tests/test_json_tables.py fuzzes it against the oracle.
"""
import os
import sys

(VALUE, AFTER, KEY_OR_CLOSE, KEY, COLON, VALUE_OR_CLOSE, STR, ESC, HEX, SUR_BS, SUR_U, MINUS, ZERO, INT, DOT, FRAC, E, ESIGN, EXP,
 T1, T2, T3, F1, F2, F3, F4, N1, N2, N3, BAD, STR_END) = range(31)
STATE_NAMES = ["VALUE", "AFTER", "KEY_OR_CLOSE", "KEY", "COLON", "VALUE_OR_CLOSE", "STR", "ESC", "HEX", "SUR_BS", "SUR_U", "MINUS", "ZERO",
               "INT", "DOT", "FRAC", "E", "ESIGN", "EXP", "T1", "T2", "T3", "F1", "F2", "F3", "F4", "N1", "N2", "N3", "BAD", "STR_END"]
(C_SPACE, C_QUOTE, C_BSLASH, C_COMMA, C_COLON, C_LBRACE, C_RBRACE, C_LBRACK, C_RBRACK, C_ZERO, C_DIGIT, C_DOT, C_e, C_E, C_PLUS, C_MINUS,
 C_t, C_r, C_u, C_a, C_l, C_s, C_f, C_n, C_WSCTL, C_OTHER, C_CTRL) = range(27)
NCLS = 27
A_NONE, A_PUSH_OBJ, A_PUSH_ARR, A_POP_OBJ, A_POP_ARR, A_COMMA, A_KEYSTR, A_VALSTR = range(8)


def byte_class(b: int) -> int:
    ch = chr(b)
    single = {' ': C_SPACE, '"': C_QUOTE, '\\': C_BSLASH, ',': C_COMMA, ':': C_COLON, '{': C_LBRACE, '}': C_RBRACE, '[': C_LBRACK,
              ']': C_RBRACK, '0': C_ZERO, '.': C_DOT, 'e': C_e, 'E': C_E, '+': C_PLUS, '-': C_MINUS, 't': C_t, 'r': C_r, 'u': C_u,
              'a': C_a, 'l': C_l, 's': C_s, 'f': C_f, 'n': C_n}
    if ch in single:
        return single[ch]
    if ch in "123456789":
        return C_DIGIT
    if ch in "\t\n\r":
        return C_WSCTL
    if b < 0x20:
        return C_CTRL
    return C_OTHER


def is_ws(c):
    return c in (C_SPACE, C_WSCTL)


def is_digit(c):
    return c in (C_ZERO, C_DIGIT)


def after(c):
    if is_ws(c):
        return AFTER, A_NONE
    if c == C_COMMA:
        return AFTER, A_COMMA
    if c == C_RBRACE:
        return AFTER, A_POP_OBJ
    if c == C_RBRACK:
        return AFTER, A_POP_ARR
    return BAD, A_NONE


def value(c):
    if is_ws(c):
        return VALUE, A_NONE
    table = {C_QUOTE: (STR, A_VALSTR), C_LBRACE: (KEY_OR_CLOSE, A_PUSH_OBJ), C_LBRACK: (VALUE_OR_CLOSE, A_PUSH_ARR), C_t: (T1, 0),
             C_f: (F1, 0), C_n: (N1, 0), C_MINUS: (MINUS, 0), C_ZERO: (ZERO, 0), C_DIGIT: (INT, 0)}
    return table.get(c, (BAD, A_NONE))


def transition(st, c):
    if st == VALUE:
        return value(c)
    if st == VALUE_OR_CLOSE:
        if is_ws(c):
            return st, 0
        if c == C_RBRACK:
            return AFTER, A_POP_ARR
        return value(c)
    if st == KEY_OR_CLOSE:
        if is_ws(c):
            return st, 0
        if c == C_RBRACE:
            return AFTER, A_POP_OBJ
        return (STR, A_KEYSTR) if c == C_QUOTE else (BAD, 0)
    if st == KEY:
        if is_ws(c):
            return st, 0
        return (STR, A_KEYSTR) if c == C_QUOTE else (BAD, 0)
    if st == COLON:
        if is_ws(c):
            return st, 0
        return (VALUE, 0) if c == C_COLON else (BAD, 0)
    if st == STR:
        if c == C_QUOTE:
            return STR_END, 0
        if c == C_BSLASH:
            return ESC, 0
        if c in (C_WSCTL, C_CTRL):
            return BAD, 0
        return STR, 0
    if st == MINUS:
        return (ZERO, 0) if c == C_ZERO else ((INT, 0) if c == C_DIGIT else (BAD, 0))
    if st == ZERO:
        if c == C_DOT:
            return DOT, 0
        if c in (C_e, C_E):
            return E, 0
        if is_digit(c):
            return BAD, 0
        return after(c)
    if st == INT:
        if is_digit(c):
            return INT, 0
        if c == C_DOT:
            return DOT, 0
        if c in (C_e, C_E):
            return E, 0
        return after(c)
    if st == DOT:
        return (FRAC, 0) if is_digit(c) else (BAD, 0)
    if st == FRAC:
        if is_digit(c):
            return FRAC, 0
        if c in (C_e, C_E):
            return E, 0
        return after(c)
    if st == E:
        if c in (C_PLUS, C_MINUS):
            return ESIGN, 0
        return (EXP, 0) if is_digit(c) else (BAD, 0)
    if st == ESIGN:
        return (EXP, 0) if is_digit(c) else (BAD, 0)
    if st == EXP:
        return (EXP, 0) if is_digit(c) else after(c)
    lit = {T1: (C_r, T2), T2: (C_u, T3), T3: (C_e, AFTER), F1: (C_a, F2), F2: (C_l, F3), F3: (C_s, F4), F4: (C_e, AFTER),
           N1: (C_u, N2), N2: (C_l, N3), N3: (C_l, AFTER)}
    if st in lit:
        want, nxt = lit[st]
        return (nxt, 0) if c == want else (BAD, 0)
    if st == AFTER:
        return after(c)
    return BAD, 0   # BAD, STR_END (never a source), and the escape states (handled outside the table)


CLS = [byte_class(b) for b in range(256)]
TRANS = [0] * (32 * NCLS)
for _st in range(31):
    for _c in range(NCLS):
        n, a = transition(_st, _c)
        TRANS[_st * NCLS + _c] = n | (a << 5)
for _c in range(NCLS):
    TRANS[31 * NCLS + _c] = BAD


def simulate(data: bytes) -> bool:
    """The device algorithm, byte for byte (json_valid_table in cell_parsers.cuh)."""
    st, depth, ctx, key = VALUE, 0, 0, False   # ctx: 0 top level, 1 object, 2 array
    lo = hi = 0                                # container bit stack, bit d = 1 when level d is an object
    aux = hexn = 0
    low_sur = False
    for b in data:
        if st == BAD:
            break
        if ESC <= st <= SUR_U:                 # rare divergent block
            ch = chr(b)
            if st == ESC:
                if ch == 'u':
                    st, hexn, aux = HEX, 4, 0
                elif ch in '"\\/bfnrt':
                    st = BAD if low_sur else STR
                else:
                    st = BAD
            elif st == HEX:
                if ch not in "0123456789abcdefABCDEF":
                    st = BAD
                else:
                    aux = aux * 16 + int(ch, 16)
                    hexn -= 1
                    if hexn == 0:
                        if low_sur:
                            st = STR if 0xDC00 <= aux <= 0xDFFF else BAD
                            low_sur = False
                        elif 0xDC00 <= aux <= 0xDFFF:
                            st = BAD
                        elif 0xD800 <= aux <= 0xDBFF:
                            st = SUR_BS
                        else:
                            st = STR
            elif st == SUR_BS:
                st = SUR_U if ch == '\\' else BAD
            else:
                if ch == 'u':
                    st, hexn, aux, low_sur = HEX, 4, 0, True
                else:
                    st = BAD
            continue
        e = TRANS[st * NCLS + CLS[b]]
        nst, act = e & 31, e >> 5
        if act in (A_PUSH_OBJ, A_PUSH_ARR):
            if depth >= 127:
                nst = BAD
            else:
                bit = 1 if act == A_PUSH_OBJ else 0
                if depth < 64:
                    lo = (lo & ~(1 << depth)) | (bit << depth)
                else:
                    hi = (hi & ~(1 << (depth - 64))) | (bit << (depth - 64))
                depth += 1
                ctx = 1 if bit else 2
        elif act in (A_POP_OBJ, A_POP_ARR):
            if ctx != (1 if act == A_POP_OBJ else 2):
                nst = BAD
            else:
                depth -= 1
                if depth == 0:
                    ctx = 0
                else:
                    p = depth - 1
                    bit = (lo >> p) & 1 if p < 64 else (hi >> (p - 64)) & 1
                    ctx = 1 if bit else 2
        elif act == A_COMMA:
            nst = BAD if ctx == 0 else (KEY if ctx == 1 else VALUE)
        elif act == A_KEYSTR:
            key, low_sur = True, False
        elif act == A_VALSTR:
            key, low_sur = False, False
        if nst == STR_END:
            nst = COLON if key else AFTER
        st = nst
    return st != BAD and depth == 0 and st in (AFTER, ZERO, INT, FRAC, EXP)


def emit(path: str):
    def rows(vals, per):
        return ",\n    ".join(", ".join(str(v) for v in vals[i:i + per]) for i in range(0, len(vals), per))
    with open(path, "w") as f:
        f.write("// GENERATED by tools/gen_json_tables.py — do not edit.\n")
        f.write("// Byte classes and transitions of the table-driven serde_json acceptor (see the generator for the grammar).\n")
        f.write("#pragma once\n#include <cstdint>\nnamespace etl {\n")
        f.write(f"constexpr int kJsonClasses = {NCLS};\n")
        f.write("enum : uint32_t { " + ", ".join(f"JT_{n} = {i}" for i, n in enumerate(STATE_NAMES)) + " };\n")
        f.write("enum : uint32_t { JA_NONE = 0, JA_PUSH_OBJ, JA_PUSH_ARR, JA_POP_OBJ, JA_POP_ARR, JA_COMMA, JA_KEYSTR, JA_VALSTR };\n")
        f.write("// [0,256): class of each byte; [256, 256 + 32*kJsonClasses): next state | action << 5\n")
        f.write("__device__ const uint8_t kJsonTables[256 + 32 * kJsonClasses] = {\n    " + rows(CLS + TRANS, 32) + "};\n")
        # the same transitions with the class lookup folded in: one load per byte instead of two dependent ones.
        # 32 states x 256 bytes; k_rows brings it into shared memory with one bulk copy per CTA.
        t2 = [TRANS[st * NCLS + CLS[b]] for st in range(32) for b in range(256)]
        f.write("// [state << 8 | byte]: next state | action << 5 (class lookup folded in)\n")
        f.write("__device__ __align__(16) const uint8_t kJsonT2[32 * 256] = {\n    " + rows(t2, 64) + "};\n")
        f.write("}  // namespace etl\n")


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "..", "etl_b200", "csrc", "json_tables.cuh")
    emit(os.path.normpath(out))
    print("wrote", os.path.normpath(out), file=sys.stderr)
