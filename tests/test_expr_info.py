"""hs_expression_info / hs_expression_ext_info: the reference's own table of expected values
(unit/hyperscan/expr_info.cpp:178-262, the rows without approximate matching) -- widths, and whether matches can
arrive out of order / at the end of the data / only there."""
import ctypes as C

import pytest

UINT_MAX = 0xffffffff


class Info(C.Structure):
    _fields_ = [("min_width", C.c_uint), ("max_width", C.c_uint), ("unordered_matches", C.c_char),
                ("matches_at_eod", C.c_char), ("matches_only_at_eod", C.c_char)]


# (pattern, ext, min, max, unordered_matches, matches_at_eod, matches_only_at_eod)
TABLE = [
    (b"abc", None, 3, 3, 0, 0, 0), (b"abc.*def", None, 6, UINT_MAX, 0, 0, 0), (b"abc|defghi", None, 3, 6, 0, 0, 0),
    (b"abc(def)?", None, 3, 6, 0, 0, 0), (b"abc(def){0,3}", None, 3, 12, 0, 0, 0), (b"abc(def){1,4}", None, 6, 15, 0, 0, 0),
    (b"", None, 0, 0, 0, 0, 0), (b"^", None, 0, 0, 0, 0, 0), (b"^\\b", None, 0, 0, 1, 0, 0), (b"\\b$", None, 0, 0, 1, 1, 1),
    (b"(?m)\\b$", None, 0, 0, 1, 1, 0), (b"\\A", None, 0, 0, 0, 0, 0), (b"\\z", None, 0, 0, 0, 1, 1), (b"\\Z", None, 0, 0, 1, 1, 1),
    (b"$", None, 0, 0, 1, 1, 1), (b"(?m)$", None, 0, 0, 1, 1, 0), (b"^foo", None, 3, 3, 0, 0, 0),
    (b"^foo.*bar", None, 6, UINT_MAX, 0, 0, 0), (b"^foo.*bar?", None, 5, UINT_MAX, 0, 0, 0),
    (b"^foo.*bar$", None, 6, UINT_MAX, 1, 1, 1), (b"^foobar$", None, 6, 6, 1, 1, 1), (b"foobar$", None, 6, 6, 1, 1, 1),
    (b"^.*foo", None, 3, UINT_MAX, 0, 0, 0), (b"foo\\b", None, 3, 3, 1, 1, 0), (b"foo.{1,13}bar", None, 7, 19, 0, 0, 0),
    (b"foo.{10,}bar", None, 16, UINT_MAX, 0, 0, 0), (b"foo.{0,10}bar", None, 6, 16, 0, 0, 0), (b"foo.{,10}bar", None, 12, 12, 0, 0, 0),
    (b"foo.{10}bar", None, 16, 16, 0, 0, 0), (b"(^|\n)foo", None, 3, 4, 0, 0, 0), (b"(^\n|)foo", None, 3, 4, 0, 0, 0),
    (b"(?m)^foo", None, 3, 3, 0, 0, 0), (b"\\bfoo", None, 3, 3, 0, 0, 0), (b"^\\bfoo", None, 3, 3, 0, 0, 0),
    (b"(?m)^\\bfoo", None, 3, 3, 0, 0, 0), (b"\\Bfoo", None, 3, 3, 0, 0, 0), (b"(foo|bar\\z)", None, 3, 3, 0, 1, 0),
    (b"(foo|bar)\\z", None, 3, 3, 0, 1, 1),
    # extended parameters
    (b"^abc.*def", {"max_offset": 10}, 6, 10, 0, 0, 0), (b"^abc.*def", {"min_length": 100}, 100, UINT_MAX, 0, 0, 0),
    (b"abc.*def", {"max_offset": 10}, 6, 10, 0, 0, 0), (b"abc.*def", {"min_length": 100}, 100, UINT_MAX, 0, 0, 0),
    (b"abc.*def", {"min_length": 5}, 6, UINT_MAX, 0, 0, 0),
]


def _info(hs, pat, ext):
    L = hs.lib()
    out = C.POINTER(Info)()
    err = C.POINTER(hs.CompileError)()
    x = None
    if ext:
        x = hs.ExprExt()
        for k, v in ext.items():
            x.flags |= {"min_offset": 1, "max_offset": 2, "min_length": 4}[k]
            setattr(x, k, v)
    L.hs_expression_ext_info.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.hs_expression_ext_info(pat, 0, C.byref(x) if x else None, C.byref(out), C.byref(err))
    assert rc == 0, (pat, err.contents.message if err else None)
    i = out.contents
    got = (i.min_width, i.max_width, ord(i.unordered_matches), ord(i.matches_at_eod), ord(i.matches_only_at_eod))
    C.CDLL(None).free(C.cast(out, C.c_void_p))
    return got


@pytest.mark.parametrize("row", TABLE, ids=[repr(r[0]) + (str(r[1]) if r[1] else "") for r in TABLE])
def test_expression_info_table(hs, row):
    pat, ext, *want = row
    assert _info(hs, pat, ext) == tuple(want)
