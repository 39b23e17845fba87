"""CPU-only checks of the C-ABI boundary: the library loads without a GPU, exports every
symbol include/lpnms.h declares, and validates its arguments before touching CUDA."""
import ctypes
import os
import re

import pytest

from yolo_lp_b200 import _abi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()           # no-op when liblpnms.so is newer than its sources
    return _abi.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lpnms.h")).read()
    return sorted(set(re.findall(r"LP_API\s+[\w\s\*]+?\b(lp_\w+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 11
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lpnms.h but not exported"
        assert n in _abi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_abi.SIGNATURES) == names


def test_version_and_errors(lib):
    assert lib.lp_version() == 200
    assert lib.lp_error_string(0) == b"ok"
    assert b"threshold" in lib.lp_error_string(-5)
    assert b"workspace" in lib.lp_error_string(-4)


def test_workspace_query(lib):
    small = _abi.nms_workspace_bytes(1, 8400, 300)
    big = _abi.nms_workspace_bytes(32, 8400, 300)
    assert 0 < small < big
    assert big >= 32 * 8400 * 8                      # one 64-bit key per anchor
    assert _abi.nms_workspace_bytes(32, 33600, 300) >= 32 * 65536 * 8   # pow2-padded for the global sort
    n = ctypes.c_size_t()
    assert lib.lp_nms_workspace_bytes(0, 8400, 300, ctypes.byref(n)) == -2
    assert lib.lp_nms_workspace_bytes(1, 8400, 300, None) == -1
    assert lib.lp_nms_workspace_bytes(1 << 20, 1 << 20, 300, ctypes.byref(n)) == -2


def test_argument_validation_without_gpu(lib):
    # every check below is rejected before any CUDA call is made
    f = lib.lp_nms_f32
    assert f(None, 1, 8, 0.25, 0.45, 300, 30000, None, 0, None, None, None, None, 0, None, None) == -1
    buf = ctypes.create_string_buffer(1 << 16)
    base = ctypes.addressof(buf)
    p = (base + 255) // 256 * 256
    args = dict(pred=p, ws=p + 4096, out=p + 32768, counts=p + 49152)
    assert f(args["pred"], 1, 8, 1.5, 0.45, 300, 30000, args["ws"], 1 << 30, args["out"], args["counts"], None, None, 0, None, None) == -5
    assert f(args["pred"], 1, 8, 0.25, -0.1, 300, 30000, args["ws"], 1 << 30, args["out"], args["counts"], None, None, 0, None, None) == -5
    assert f(args["pred"] + 8, 1, 8, 0.25, 0.45, 300, 30000, args["ws"], 1 << 30, args["out"], args["counts"], None, None, 0, None, None) == -3
    assert f(args["pred"], 1, 8, 0.25, 0.45, 300, 30000, args["ws"], 16, args["out"], args["counts"], None, None, 0, None, None) == -4
    assert f(args["pred"], 0, 8, 0.25, 0.45, 300, 30000, args["ws"], 1 << 30, args["out"], args["counts"], None, None, 0, None, None) == -2
    assert lib.lp_rescale_f32(p, 4, 8, 0.0, 0.0, 1.0, 10.0, 10.0, 0, None) == -2     # row stride < 12
    assert lib.lp_rescale_f32(p, 4, 12, 0.0, 0.0, 0.0, 10.0, 10.0, 0, None) == -6    # ratio must be > 0
    assert lib.lp_dist2bbox_f32(p + 4, p, 1, 8, 0, p, None) == -3
    assert lib.lp_detect_decode_f32(None, 3, 1, p, None, None) == -1


def test_check_maps_errors_like_the_reference(lib):
    with pytest.raises(AssertionError):
        _abi.check("lp_nms_f32", -5)          # nms.py:57-58 raise AssertionError
    with pytest.raises(ValueError):
        _abi.check("lp_nms_f32", -2)
    with pytest.raises(_abi.LpError):
        _abi.check("lp_nms_f32", 700)
    _abi.check("lp_nms_f32", 0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "yolo_lp_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
            assert "/root/reference" not in src, fn


def test_ctypes_signatures_have_the_headers_parameter_counts(lib):
    """ABI drift guard: every prototype in include/lpnms.h and its ctypes entry must agree on the
    number of parameters (a missing / extra argument would otherwise only show as a crash)."""
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lpnms.h")).read(), flags=re.S)
    protos = re.findall(r"LP_API\s+[\w\s\*]+?\b(lp_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(protos) == len(_abi.SIGNATURES)
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else len([q for q in params.split(",") if q.strip()])
        assert n == len(_abi.SIGNATURES[name][1]), f"{name}: header has {n} parameters, ctypes table {len(_abi.SIGNATURES[name][1])}"


def test_f16_entries_validate_like_the_f32_ones(lib):
    for name in ("lp_nms_f16", "lp_nms_f32"):
        f = getattr(lib, name)
        assert f(None, 1, 8, 0.25, 0.45, 300, 30000, None, 0, None, None, None, None, 0, None, None) == -1
        buf = ctypes.create_string_buffer(1 << 16)
        p = (ctypes.addressof(buf) + 255) // 256 * 256
        assert f(p, 1, 8, 1.5, 0.45, 300, 30000, p + 4096, 1 << 30, p + 32768, p + 49152, None, None, 0, None, None) == -5
        assert f(p + 8, 1, 8, 0.25, 0.45, 300, 30000, p + 4096, 1 << 30, p + 32768, p + 49152, None, None, 0, None, None) == -3
        assert f(p, 1, 8, 0.25, 0.45, 300, 30000, p + 4096, 16, p + 32768, p + 49152, None, None, 0, None, None) == -4


def test_ctypes_argument_kinds_match_the_header(lib):
    """... and on the KIND of every parameter: pointer, int, long long, size_t, float or double."""
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lpnms.h")).read(), flags=re.S)
    protos = re.findall(r"LP_API\s+[\w\s\*]+?\b(lp_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)

    def header_kind(decl):
        d = " ".join(decl.split())
        if "*" in d or "lp_stream_t" in d:
            return "ptr"
        for key, kind in (("size_t", "size_t"), ("long long", "longlong"), ("double", "double"), ("float", "float"), ("int", "int")):
            if re.search(r"\b" + key + r"\b", d):
                return kind
        raise AssertionError(f"unparsed parameter {decl!r}")

    def ctypes_kind(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or isinstance(t, type(ctypes.POINTER(ctypes.c_int))):
            return "ptr"
        return {ctypes.c_int: "int", ctypes.c_double: "double", ctypes.c_float: "float", ctypes.c_size_t: "size_t",
                ctypes.c_longlong: "longlong"}[t]

    for name, params in protos:
        decls = [q for q in params.split(",") if q.strip() and q.strip() != "void"]
        want = [header_kind(d) for d in decls]
        got = [ctypes_kind(t) for t in _abi.SIGNATURES[name][1]]
        # size_t and c_size_t / c_ulong are the same thing on this ABI
        assert got == want, f"{name}: header {want} vs ctypes {got}"


def test_pipelined_entries_validate_before_queueing(lib):
    """lp_nms_pipelined_* / lp_detect_pipelined_* must reject every bad K2 argument before
    K1 / KF is queued (a half-queued step would leave the workspace counters dirty).  Without a GPU the
    proof is that the error comes back as an LP_E_* code, not as a CUDA error from a stream call."""
    buf = ctypes.create_string_buffer(1 << 16)
    p = (ctypes.addressof(buf) + 255) // 256 * 256
    ev = 1  # a non-NULL event handle; never dereferenced because validation fails first
    for name in ("lp_nms_pipelined_f32", "lp_nms_pipelined_f16"):
        f = getattr(lib, name)
        ok = dict(pred=p, B=1, A=8, conf=0.25, iou=0.45, max_det=300, max_nms=30000, ws=p + 4096, ws_bytes=1 << 30,
                  out=p + 32768, counts=p + 49152)

        def call(**kw):
            a = dict(ok, **kw)
            return f(a["pred"], a["B"], a["A"], a["conf"], a["iou"], a["max_det"], a["max_nms"], a["ws"], a["ws_bytes"],
                     a["out"], a["counts"], None, None, 0, None, None, None, ev, None, None, None, None)
        assert call(counts=None) == -1          # K2 argument
        assert call(out=None) == -1
        assert call(max_nms=0) == -2
        assert call(iou=1.5) == -5
        assert call(conf=-0.5) == -5
        assert call(ws_bytes=64) == -4
        assert call(pred=p + 8) == -3
    lv = (_abi.LpLevel * 1)()
    for g in range(8):
        lv[0].cls[g] = p
    lv[0].reg, lv[0].cor, lv[0].h, lv[0].w, lv[0].stride = p, p, 2, 4, 8.0
    for name in ("lp_detect_pipelined_f32", "lp_detect_pipelined_f16"):
        f = getattr(lib, name)

        def call(iou=0.45, max_nms=30000, counts=p + 49152, ws_bytes=1 << 30):
            return f(lv, 1, 1, 0.25, iou, 300, max_nms, p + 4096, ws_bytes, p + 32768, counts, None, None, 0, None, None,
                     None, ev, None, None, None, None)
        assert call(counts=None) == -1
        assert call(max_nms=0) == -2
        assert call(iou=2.0) == -5
        assert call(ws_bytes=64) == -4


def test_opts_struct_matches_the_header():
    text = open(os.path.join(ROOT, "include", "lpnms.h")).read()
    body = re.search(r"typedef struct lp_opts \{(.*?)\} lp_opts_t;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [d.split()[-1].lstrip("*") for d in body.split(";") if d.strip()]
    assert fields == [n for n, _ in _abi.LpOpts._fields_]
    o = _abi.opts(filter_ctas=7, no_tma=True, timing=4096)
    assert (o.filter_ctas, o.no_tma, o.timing) == (7, 1, 4096)
