"""TEST INFRASTRUCTURE. Builds tests/cpp/_build/troyn_dropin_tests: a gtest binary that runs

  1. the reference's OWN CUDA tests for the hot path (selected TEST bodies of <reference>/test/evaluator_cuda.cu,
     compiled from where they lie: the selected bodies are assembled into a scratch file under /tmp at build time,
     nothing is copied into this repository) against include/troyn/troy_cuda.cuh + libtroy_b200.so, and
  2. tests/cpp/dropin_parity.cpp (word-for-word parity with the reference's CPU evaluator through the C++ layer).

The binary links libtroy_b200.so and oracle/_ref/libtroyref.so (the reference's unmodified host classes, which the
drop-in layer reuses for key generation / encryption / encoding exactly as the reference's CUDA classes do). It is
git-ignored but travels to the GPU box. Needs /root/reference at BUILD time only.
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("TROY_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(ROOT, "tests", "cpp", "_build")
OUT = os.path.join(OUT_DIR, "troyn_dropin_tests")

# hot-path tests of the reference's CUDA suite
REFERENCE_TESTS = [
    "BFVEncryptNegateDecrypt", "BFVEncryptAddDecrypt", "BGVEncryptNegateDecrypt", "BGVEncryptAddDecrypt", "CKKSEncryptAddDecrypt",
    "BFVEncryptSubDecrypt", "BGVEncryptSubDecrypt",
    "BFVEncryptMultiplyDecrypt", "BGVEncryptMultiplyDecrypt", "BFVRelinearize", "BGVRelinearize",
    "CKKSEncryptMultiplyRelinDecrypt", "CKKSEncryptSquareRelinDecrypt", "CKKSEncryptMultiplyRelinRescaleDecrypt",
    "CKKSEncryptSquareRelinRescaleDecrypt", "CKKSEncryptModSwitchDecrypt",
    "CKKSEncryptRotateDecrypt", "CKKSEncryptRescaleRotateDecrypt", "BFVEncryptSquareDecrypt", "BGVEncryptSquareDecrypt",
    "TransformEncryptedToFromNTT", "BFVEncryptApplyGaloisDecrypt", "BFVEncryptRotateMatrixDecrypt",
    "BGVEncryptApplyGaloisDecrypt", "BGVEncryptRotateMatrixDecrypt", "BFVEncryptModSwitchToNextDecrypt", "BGVEncryptModSwitchToNextDecrypt",
    # ciphertext (x) plaintext
    "CKKSEncryptAddPlainDecrypt", "CKKSEncryptSubPlainDecrypt", "BFVEncryptAddPlainDecrypt", "BFVEncryptSubPlainDecrypt",
    "BFVEncryptMultiplyPlainDecrypt", "BGVEncryptAddPlainDecrypt", "BGVEncryptSubPlainDecrypt", "BGVEncryptMultiplyPlainDecrypt",
    "CKKSEncryptMultiplyByNumberDecrypt", "TransformPlainToNTT", "BFVEncryptMultiplyPlainNTTDecrypt", "BGVEncryptMultiplyPlainNTTDecrypt",
    # compositions over the path: product trees, powers, sums, key switching to another secret key, multi-level switches, wire format
    "BFVEncryptMultiplyManyDecrypt", "BGVEncryptMultiplyManyDecrypt", "BFVEncryptExponentiateDecrypt", "BGVEncryptExponentiateDecrypt",
    "BFVEncryptAddManyDecrypt", "BGVEncryptAddManyDecrypt", "BFVKeySwitching", "BFVEncryptModSwitchToDecrypt", "BGVEncryptModSwitchToDecrypt",
    "CKKSEncryptNaiveMultiplyDecrypt", "CKKSEncryptMultiplyRelinRescaleModSwitchAddDecrypt", "SerializeTest",
    # LWE extraction, assembly, field trace and packing
    "BFVAssembleLWETest",
]


def extract(text, names):
    """TEST bodies by name; a body runs to the next TEST at the same indentation (or the namespace's closing brace)."""
    starts = [(m.start(), m.group(1)) for m in re.finditer(r"^    TEST\(EvaluatorCudaTest, (\w+)\)", text, re.M)]
    end_ns = text.rstrip().rfind("}")
    out, found = [], []
    for i, (pos, name) in enumerate(starts):
        if name in names:
            stop = starts[i + 1][0] if i + 1 < len(starts) else end_ns
            out.append(text[pos:stop])
            found.append(name)
    return "".join(out), found


def main():
    src = os.path.join(REF, "test", "evaluator_cuda.cu")
    if not os.path.exists(src):
        print(f"reference sources not present at {REF}; keeping prebuilt {OUT}")
        return 0
    os.makedirs(OUT_DIR, exist_ok=True)
    text = open(src).read()
    bodies, found = extract(text, set(REFERENCE_TESTS))
    missing = sorted(set(REFERENCE_TESTS) - set(found))
    if missing:
        print("note: not present in the reference test file:", ", ".join(missing))
    head = text[:text.index("namespace troytest")]
    aliases = "\n".join(l for l in head.splitlines() if l.startswith("using "))
    gt = os.path.join(REF, "extern", "googletest", "googletest")
    with tempfile.TemporaryDirectory(prefix="troyn_dropin_") as tmp:
        gen = os.path.join(tmp, "reference_cuda_tests.cpp")
        with open(gen, "w") as f:
            f.write('#include "troy_cuda.cuh"\n#include <cstddef>\n#include <cstdint>\n#include <ctime>\n#include <string>\n'
                    '#include "gtest/gtest.h"\n' + aliases + "\nnamespace troytest\n{\n" + bodies + "}\n")
        # the reference's application header for the HE matmul, minus the line that includes the reference's own CUDA API
        app = open(os.path.join(REF, "app", "LinearHelperCKKS.cuh")).read()
        with open(os.path.join(tmp, "LinearHelperCKKS.cuh"), "w") as f:
            f.write("\n".join(l for l in app.splitlines() if "troy_cuda.cuh" not in l) + "\n")
        # the reference's encryptor and CKKS-encoder CUDA test files, whole: only the line that includes the reference's own CUDA API
        # is redirected to the drop-in header (scratch copies under /tmp at build time, nothing enters the repository)
        whole = []
        for name in ("encryptor_cuda.cu", "ckks_cuda.cu"):
            text_w = open(os.path.join(REF, "test", name)).read().replace('#include "../src/troy_cuda.cuh"', '#include "troy_cuda.cuh"')
            dst = os.path.join(tmp, "reference_" + name.replace(".cu", ".cpp"))
            with open(dst, "w") as f:
                f.write(text_w)
            whole.append(dst)
        inc = ["-I" + tmp, "-I" + os.path.join(ROOT, "include", "troyn"), "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(REF, "src"),
               "-I" + os.path.join(gt, "include"), "-I" + gt]
        objs = []
        jobs = [(gen, "ref_tests.o"), (os.path.join(ROOT, "tests", "cpp", "dropin_parity.cpp"), "parity.o"),
                (os.path.join(ROOT, "tests", "cpp", "dropin_linear.cpp"), "linear.o"),
                (os.path.join(gt, "src", "gtest-all.cc"), "gtest-all.o"), (os.path.join(gt, "src", "gtest_main.cc"), "gtest_main.o")]
        jobs += [(w, os.path.basename(w) + ".o") for w in whole]
        procs = []
        for s, o in jobs:
            o = os.path.join(tmp, o)
            objs.append(o)
            procs.append(subprocess.Popen(["g++", "-std=c++17", "-O1", "-w", *inc, "-c", s, "-o", o]))
        if any(p.wait() for p in procs):
            return 1
        cmd = ["g++", "-o", OUT, *objs, "-L" + os.path.join(ROOT, "troy_b200"), "-ltroy_b200", "-L" + os.path.join(ROOT, "oracle", "_ref"),
               "-ltroyref", "-Wl,-rpath,$ORIGIN/../../../troy_b200", "-Wl,-rpath,$ORIGIN/../../../oracle/_ref", "-lpthread"]
        print("+", " ".join(cmd), flush=True)
        if subprocess.call(cmd):
            return 1
        # the reference's own timing program against the drop-in (the other build, against the reference's CUDA library, is
        # oracle/_ref/timetest_reference)
        tt = os.path.join(tmp, "reference_timetest.cpp")
        with open(tt, "w") as f:
            f.write(open(os.path.join(REF, "test", "timetest.cu")).read().replace('#include "../src/troy_cuda.cuh"', '#include "troy_cuda.cuh"'))
        cmd = ["g++", "-std=c++17", "-O2", "-w", *inc, tt, "-o", os.path.join(OUT_DIR, "timetest_dropin"), "-L" + os.path.join(ROOT, "troy_b200"),
               "-ltroy_b200", "-L" + os.path.join(ROOT, "oracle", "_ref"), "-ltroyref", "-Wl,-rpath,$ORIGIN/../../../troy_b200",
               "-Wl,-rpath,$ORIGIN/../../../oracle/_ref", "-lpthread"]
        if subprocess.call(cmd):
            print("note: the reference's timetest.cu did not build against the drop-in")
    print(f"built {OUT}: {len(found)} reference evaluator tests + the reference's encryptor_cuda.cu and ckks_cuda.cu + DropInParity + DropInLinear")
    return 0


if __name__ == "__main__":
    sys.exit(main())
