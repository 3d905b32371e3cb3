"""CPU check of DEVICE code: the token lexer of the text -> HBM parse kernel (text_convert_token in
pytuq_b200/csrc/pcq_text.cu) is plain C++ apart from its qualifiers, so its source is lifted out of the .cu file,
compiled for the host with g++ and run against the shared lexer of the host library (pcq_parse_decimal_fast,
csrc/pcq_decimal.h) on random tokens: whatever the device lexer accepts, the shared one accepts with the same bits
(it may decline more -- those tokens go to the host fix-up), and on well-formed tokens both accept equally often."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "pytuq_b200", "csrc")

MAIN = r'''
static uint64_t st = 88172645463325252ull;
static uint64_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; }
int main() {
    long bad = 0, acc = 0, acc_shared = 0, n = 0;
    char buf[128];
    for (long it = 0; it < 3000000; ++it) {
        int len;
        const uint64_t r = rnd();
        double v; uint64_t u = rnd(); memcpy(&v, &u, 8);
        switch (r % 8) {
            case 0: len = snprintf(buf, sizeof buf, "%.18e", v); break;
            case 1: len = snprintf(buf, sizeof buf, "%.*e", (int)(rnd() % 22), v); break;
            case 2: len = snprintf(buf, sizeof buf, "%.*f", (int)(rnd() % 12), std::fmod(v, 1e9)); break;
            case 3: len = snprintf(buf, sizeof buf, "%lld", (long long)(rnd() % 100000000000ull) - 50000000000ll); break;
            case 4: len = snprintf(buf, sizeof buf, "%s%llu.%llue%s%d", (r & 16) ? "-" : (r & 32) ? "+" : "",
                                   (unsigned long long)(rnd() % 1000), (unsigned long long)(rnd() % 1000000),
                                   (r & 64) ? "-" : (r & 128) ? "+" : "", (int)(rnd() % 400)); break;
            case 5: len = snprintf(buf, sizeof buf, "000%g", (double)(int)(rnd() % 100000) / 64.0); break;
            case 6: { const char* al = "0123456789+-.eE"; len = 1 + rnd() % 12;
                      for (int k = 0; k < len; ++k) buf[k] = al[rnd() % 15]; buf[len] = 0; break; }
            default: len = snprintf(buf, sizeof buf, ".%llu", (unsigned long long)(rnd() % 100000));
        }
        if (len <= 0 || len > 100 || !std::isfinite(v)) continue;
        ++n;
        double x = 0, y = 0;
        const bool ok1 = text_convert_token((const unsigned char*)buf, len, &x);
        const bool ok2 = pcq_parse_decimal_fast(buf, (size_t)len, &y, PCQIO_POW10);
        acc += ok1; acc_shared += ok2;
        if (ok1 && (!ok2 || memcmp(&x, &y, 8) != 0)) { if (bad < 10) printf("MISMATCH '%s'\n", buf); ++bad; }
    }
    printf("%ld %ld %ld %ld\n", n, acc, acc_shared, bad);
    return 0;
}
'''


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs a host C++ compiler")
def test_device_token_lexer_agrees_with_the_shared_lexer(tmp_path):
    src = open(os.path.join(CSRC, "pcq_text.cu")).read()
    a = src.index("__device__ __forceinline__ bool text_convert_token(")
    b = src.index("__global__", a)
    fn = src[a:b].replace("__device__ __forceinline__ ", "static ").replace("__restrict__", "").replace("d_pow10", "PCQIO_POW10")
    cpp = tmp_path / "lexer.cpp"
    cpp.write_text("#include <cmath>\n#include <cstdio>\n#include <cstdlib>\n#include <cstring>\n"
                   "#include \"pcq_io_pow10.h\"\n#include \"pcq_decimal.h\"\n" + fn + MAIN)
    exe = tmp_path / "lexer"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", CSRC, str(cpp), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    n, acc, acc_shared, bad = map(int, out[-1].split())
    assert n > 2_500_000 and bad == 0, out
    assert acc == acc_shared and acc > 0.8 * n, out
