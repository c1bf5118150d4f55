"""The Java / JNI shim ships as SOURCE (java/, jni/): no JDK exists in this image, so these tests check what can be
checked without one -- the Java files parse (with the same Java-subset parser that executes the reference's sources),
implement the reference's interfaces method for method, every `static native` has its JNI function, mcpjni.c compiles
against a JNI header stub, and a plain C program links libmcp.so and runs the stored-document known answer."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "java", "com", "mapr", "db", "data")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _parse(name):
    import javamini
    src = open(os.path.join(JAVA, name)).read()
    return javamini.Parser(src, name).parse_unit()[0]


def test_java_sources_parse_and_mirror_the_reference_interfaces():
    codec = _parse("NativeIntegerCodec.java")
    assert "IntegerCODEC" in codec["implements"]
    sigs = {(m["name"], tuple(p[0] for p in m["params"])) for m in codec["methods"]}
    five = ("int[]", "IntWrapper", "int", "int[]", "IntWrapper")
    assert ("compress", five) in sigs and ("uncompress", five) in sigs                 # IntegerCODEC.java:36-58
    assert ("getCompressedInstruments", ("int[]",)) in sigs and ("getDecompressedInstruments", ("byte[]", "int")) in sigs
    sk = _parse("NativeSkippableCodec.java")
    assert "SkippableIntegerCODEC" in sk["implements"]
    sigs = {(m["name"], tuple(p[0] for p in m["params"])) for m in sk["methods"]}
    assert ("headlessCompress", five) in sigs and ("headlessUncompress", five + ("int",)) in sigs   # SkippableIntegerCODEC.java:43-66
    sim = _parse("SimulatePortfolios.java")
    assert any(m["name"] == "main" and "static" in m["mods"] for m in sim["methods"])
    assert any(f["name"] == "SIM_RES_TABLE_PATH" for f in sim["fields"])             # the table ManageMapRDBTable.java:15 creates


def test_every_static_native_has_its_jni_function():
    natives = [m["name"] for m in _parse("McpNative.java")["methods"] if "native" in m["mods"]]
    assert len(natives) >= 10
    c = open(os.path.join(ROOT, "jni", "mcpjni.c")).read()
    for n in natives:
        assert re.search(r"Java_com_mapr_db_data_McpNative_%s\s*\(" % n, c), f"mcpjni.c lacks the JNI function of McpNative.{n}"
    # and every C entry point the JNI layer calls is declared in include/mcp.h
    hdr = open(os.path.join(ROOT, "include", "mcp.h")).read()
    for fn in set(re.findall(r"\b(mcp_[a-z_0-9]+)\s*\(", c)) | set(re.findall(r",\s*(mcp_codec_[a-z_]+),", c)):
        assert re.search(r"\b%s\s*\(" % fn, hdr), fn


def test_mcpjni_compiles_against_a_jni_header_stub():
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "jni_stub"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "jni", "mcpjni.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _build_abi_smoke(tmp_path):
    import mcp_b200
    from mcp_b200 import native as N
    N.lib()
    libdir = os.path.dirname(N.LIB_PATH)
    exe = str(tmp_path / "abi_smoke")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_smoke.c"),
                        "-L" + libdir, "-lmcp", "-lm", "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_plain_c_program_links_the_library(tmp_path):
    """Without a GPU the program must report 'no GPU' (exit 77): mcp_create fails loudly, nothing falls back."""
    exe = _build_abi_smoke(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode in (0, 77), (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_plain_c_program_runs_the_known_answer_on_the_gpu(tmp_path):
    exe = _build_abi_smoke(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "abi_smoke ok" in r.stdout
