"""The reference-side binding of INTEGRATION.md compiled for real: oracle/_ref/kaiju-gpu = the reference's unmodified main() (kaiju.cpp),
loaders (readFMI, parseNodesDmp) and queue, linked with kaiju_b200/integration/consumer_gpu.cpp (ConsumerThread::doWork over the C ABI)
and libkaijub200.so.  Its output must equal the stock binary's."""
import gzip, os, shutil, subprocess
import pytest
from helpers import REF_DIR, have_ref

BIN = os.path.join(REF_DIR, "kaiju-gpu")


def _plain(src, dst):
    with gzip.open(src, "rb") as f, open(dst, "wb") as g:
        shutil.copyfileobj(f, g)
    return dst


def test_binding_links_and_fails_loudly_without_a_gpu(golden, tmp_path):
    """CPU part: the binary exists (built by oracle/Makefile against the reference's real structs) and, without a GPU, stops with the
    library's "no CUDA device" error instead of falling back to anything."""
    if not (have_ref() and os.path.exists(BIN)):
        pytest.skip("oracle/_ref/kaiju-gpu not built")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    fq = _plain(os.path.join(os.path.dirname(golden.fmi), "se100.fq.gz"), str(tmp_path / "se.fq"))
    p = subprocess.run([BIN, "-t", golden.nodes, "-f", golden.fmi, "-i", fq, "-a", "mem", "-z", "1", "-o", str(tmp_path / "o.tsv")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 and b"no CUDA device" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("args", [["-a", "mem"], ["-a", "mem", "-m", "5", "-X"], ["-a", "greedy", "-e", "3", "-s", "65"], ["-a", "greedy", "-e", "5"]])
@pytest.mark.parametrize("tag", ["pe150", "se100"])
def test_binding_output_equals_stock_binary(golden, tmp_path, args, tag):
    if not (have_ref() and os.path.exists(BIN)):
        pytest.skip("oracle/_ref/kaiju-gpu not built")
    g = os.path.dirname(golden.fmi)
    if tag == "pe150":
        inp = ["-i", _plain(g + "/pe150_1.fq.gz", str(tmp_path / "a.fq")), "-j", _plain(g + "/pe150_2.fq.gz", str(tmp_path / "b.fq"))]
    else:
        inp = ["-i", _plain(g + "/se100.fq.gz", str(tmp_path / "a.fq"))]
    outs = {}
    for name, exe, z in (("ref", os.path.join(REF_DIR, "kaiju"), "1"), ("gpu", BIN, "1"), ("gpu4", BIN, "4")):
        o = str(tmp_path / (name + ".tsv"))
        subprocess.run([exe, "-t", golden.nodes, "-f", golden.fmi] + inp + args + ["-z", z, "-o", o], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        outs[name] = open(o).read()
    assert outs["gpu"] == outs["ref"]                                   # -z 1: same lines in the same order
    assert sorted(outs["gpu4"].split("\n")) == sorted(outs["ref"].split("\n"))
    assert outs["ref"].count("\nC\t") > 100
