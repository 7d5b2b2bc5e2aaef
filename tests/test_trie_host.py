"""Host-side trie construction (pygsti_b200/csrc/trie_host.h): compile the invariant checker with g++ and run it."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_build_trie_invariants(tmp_path):
    exe = str(tmp_path / "trie_host_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "trie_host_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "build_trie OK" in r.stdout, r.stdout + r.stderr
