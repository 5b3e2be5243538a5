"""The C++ host mirror (nshogi-engine_b200/host): infer::B200 behind the reference's Infer
interface, the pinned multi-slot LeafPipeline, the EvalCache restatement and the move-index adaptor."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "nshogi-engine_b200", "host")


@pytest.fixture(scope="module")
def built(pkg):
    subprocess.check_call(["make", "-C", HOST, "-s", "all"])
    return HOST


def test_host_unit_cpu(built):
    """EvalCacheB200 == behaviour of reference src/mcts/evalcache.cc (164-move cap, refresh-only
    duplicates, LRU eviction in bundles of 3, feed() of CSR rows); move-index range / mirroring."""
    out = subprocess.run([os.path.join(built, "nsb_host_unit")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_unit ok" in out.stdout


def test_host_bench_fails_loudly_without_gpu(built, nb):
    if nb.device_count() > 0:
        pytest.skip("GPU present")
    out = subprocess.run([os.path.join(built, "nsb_host_bench")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 2 and "no CPU fallback" in out.stderr


def test_host_headers_compile_against_reference_interface(built):
    """infer_b200.h must also compile against the REFERENCE's own infer.h (not only our shim) when
    the reference tree is present: that is the drop-in claim of INTEGRATION.md."""
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present on this box")
    src = '#include "infer_b200.h"\nint main() { return sizeof(nshogi::engine::infer::B200) > 0 ? 0 : 1; }\n'
    cmd = ["g++", "-std=c++20", "-fsyntax-only", "-x", "c++", "-", f"-I{ref}", f"-I{HOST}",
           f"-I{os.path.join(ROOT, 'include')}", f"-I{os.path.join(HOST, 'shim')}"]
    out = subprocess.run(cmd, input=src, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr


@pytest.mark.gpu
def test_host_bench_selfcheck_gpu(built):
    """batchsize.cc-shaped run through infer::B200 (computeBlocking) and through LeafPipeline (4 pinned
    slots, positions in, fused decode out, EvalCache feed); the two paths must agree."""
    out = subprocess.run([os.path.join(built, "nsb_host_bench"), "--selfcheck", "--repeat", "50", "--blocks", "2"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    assert rec["ok"] and rec["rows_identical"] and rec["cache_hit"]
    assert rec["max_prob_diff"] < 1e-5 and rec["legal_moves"] >= 28
    assert rec["pipeline_evals_per_s"] > 0 and rec["infer_blocking_evals_per_s"] > 0


def test_selfplay_sim_fails_loudly_without_gpu(built, nb):
    if nb.device_count() > 0:
        pytest.skip("GPU present")
    out = subprocess.run([os.path.join(built, "nsb_selfplay_sim"), "--seconds", "0.1"], capture_output=True, text=True,
                         timeout=60)
    assert out.returncode == 2 and "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_selfplay_sim_gpu(built):
    """Self-play loop (frame pool -> search workers -> pinned multi-slot evaluation worker -> back) on a
    small net: every frame keeps cycling, batches fill up, no NaN rows, records are produced."""
    out = subprocess.run([os.path.join(built, "nsb_selfplay_sim"), "--channels", "128", "--blocks", "2", "--batch-size", "128",
                          "--frame-pool-size", "512", "--num-search-workers", "2", "--num-playouts", "8",
                          "--seconds", "1.0", "--warmup", "0.3"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    assert rec["nan_rows"] == 0 and rec["evals"] > 10000 and rec["records"] > 1000 and rec["games"] > 0
    assert 1 <= rec["avg_batch"] <= 128
    # a move is played every num_playouts (full search) or num_playouts / 4 (reduced search) evaluations
    assert 2.0 <= rec["evals"] / rec["records"] <= 8.0


@pytest.mark.gpu
def test_selfplay_sim_with_device_cache_gpu(built):
    """The same loop behind the device-resident evaluation cache: half of the descents revisit a recent
    leaf, so about half of the evaluations must be served from the cache."""
    out = subprocess.run([os.path.join(built, "nsb_selfplay_sim"), "--channels", "128", "--blocks", "2", "--batch-size", "128",
                          "--frame-pool-size", "512", "--num-search-workers", "2", "--num-playouts", "8", "--cache-mb", "64",
                          "--revisit-ratio", "0.5", "--seconds", "1.0", "--warmup", "0.3"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    assert rec["nan_rows"] == 0 and rec["evals"] > 10000 and rec["records"] > 1000
    assert 0.28 <= rec["cache_hit_rate"] <= 0.6, rec["cache_hit_rate"]


def test_leaf_queue_protocol_under_thread_sanitizer(built, tmp_path):
    """leaf_queue.h (lock-free in-place batch assembly, replaces reference src/mcts/evaluationqueue.cc + getBatch):
    6 filling threads, 60,000 leaves through 3 slots of 64 rows on plain memory - every leaf comes back exactly once
    with its handle, hash and CSR span - built with -fsanitize=thread (the reference's own race-detection practice,
    SURVEY.md §5), which must stay silent."""
    inc = ["-I" + HOST, "-I" + os.path.join(HOST, "shim"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(HOST, "shim")]
    exe = str(tmp_path / "unit_tsan")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", *inc, "-o", exe,
                        os.path.join(HOST, "host_unit.cc"), "-lpthread"], capture_output=True, text=True, timeout=300)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe, "--queue-stress", "6", "60000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "60000 leaves fed" in out.stdout and ": ok" in out.stdout
    assert "ThreadSanitizer" not in out.stderr
    out = subprocess.run([os.path.join(built, "nsb_host_unit"), "--queue-stress", "8", "300000"], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0 and ": ok" in out.stdout
    # the lock-free search tree (mcts_search.h: relaxed counters, claim by compare-and-swap, edges published by a release
    # store) under the same sanitizer: 6 threads grow one tree of real shogi positions, then every invariant is checked
    out = subprocess.run([exe, "--tree-stress", "6", "6000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and ": ok" in out.stdout, out.stdout + out.stderr
    assert "ThreadSanitizer" not in out.stderr
    out = subprocess.run([os.path.join(built, "nsb_host_unit"), "--tree-stress", "8", "200000"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and ": ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_infer_contract_on_plain_buffers_is_page_locked_by_the_executor(built):
    """The unmodified Infer contract with buffers the executor has never seen (plain page-aligned malloc, as an
    Evaluator built without CUDA_ENABLED would hand over): infer::B200 page-locks them on first sight, a one-slot
    executor then runs direct I/O on them; same self-check as with nsb_host_alloc memory, and faster than staged."""
    rates = {}
    for flag in ([], ["--malloc-buffers"]):
        out = subprocess.run([os.path.join(built, "nsb_host_bench"), "--selfcheck", "--slots", "1", "--repeat", "200", *flag],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        line = json.loads(out.stdout.strip().splitlines()[-1])
        assert line["ok"] and line["io_mode"] == "direct"
        rates[bool(flag)] = line["infer_blocking_evals_per_s"]
    assert rates[True] > 0.9 * rates[False]      # page-locked by the executor == allocated page-locked


def test_pipelined_evaluation_worker_inside_the_reference_worker_contract(built, tmp_path):
    """host/evaluation_worker_b200.h derives from worker::Worker.  Built here against the REFERENCE's own
    src/worker/worker.{h,cc} (compiled in place, nothing copied) and run on plain memory: three start / stop / await
    cycles, after each of which every task pushed so far has been filled once and delivered once - a stopped worker is a
    drained worker.  (The in-tree build runs the same check against this repo's stand-in, host/shim/worker/worker.h.)"""
    out = subprocess.run([os.path.join(built, "nsb_host_unit"), "--worker-cycles", "20000"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "cycles: ok" in out.stdout, out.stdout + out.stderr
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present on this box")
    exe = str(tmp_path / "unit_refworker")
    r = subprocess.run(["g++", "-std=c++20", "-O2", f"-I{ref}", f"-I{HOST}", f"-I{os.path.join(HOST, 'shim')}",
                        f"-I{os.path.join(ROOT, 'include')}", "-o", exe, os.path.join(HOST, "host_unit.cc"),
                        os.path.join(ref, "worker", "worker.cc"), "-lpthread"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe, "--worker-cycles", "20000"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "cycles: ok" in out.stdout, out.stdout + out.stderr
    # the whole self-play harness on the reference's Worker: it must wind down (search workers never run dry on their own)
    out = subprocess.run([exe, "--selfplay-loop", "3", "96", "500"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "all frames accounted for: ok" in out.stdout, out.stdout + out.stderr


def test_selfplay_and_usi_harnesses_start_play_and_wind_down_cpu(built, tmp_path):
    """host/selfplay_workers.h - the search workers, the pipelined evaluation worker and the save worker of
    nsb_selfplay_real, started and stopped exactly as its main() does - on a mock pipeline that invents the evaluations:
    real rules, real trees, real teacher records.  The run must END within the timeout (a worker::Worker is only
    stopped while its doTask() reports idle; a pool of games never goes idle by itself), with every frame back in the
    search queue exactly once and every finished game saved.  Also under -fsanitize=thread."""
    out = subprocess.run([os.path.join(built, "nsb_host_unit"), "--selfplay-loop", "4", "128", "800"], capture_output=True,
                         text=True, timeout=60)
    assert out.returncode == 0 and "all frames accounted for: ok" in out.stdout, out.stdout + out.stderr
    inc = ["-I" + HOST, "-I" + os.path.join(HOST, "shim"), "-I" + os.path.join(ROOT, "include")]
    exe = str(tmp_path / "loop_tsan")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", *inc, "-o", exe,
                        os.path.join(HOST, "host_unit.cc"), "-lpthread"], capture_output=True, text=True, timeout=300)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe, "--selfplay-loop", "3", "48", "400"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "all frames accounted for: ok" in out.stdout, out.stdout + out.stderr
    assert "ThreadSanitizer" not in out.stderr, out.stderr[-3000:]
    # the USI-style search (host/usi_search.h) on the mock pipeline: search threads filling batches in place, the
    # evaluation thread submitting, feeding and collecting leaves itself; must end, tree invariants hold, TSAN silent
    for args in (["3", "400"], ["1", "300"], ["2", "300", "nohelp"]):
        out = subprocess.run([exe, "--usi-loop", *args], capture_output=True, text=True, timeout=180)
        assert out.returncode == 0 and "invariants hold: ok" in out.stdout, out.stdout + out.stderr
        assert "ThreadSanitizer" not in out.stderr, out.stderr[-3000:]
    out = subprocess.run([os.path.join(built, "nsb_host_unit"), "--usi-loop", "6", "1500"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "invariants hold: ok" in out.stdout, out.stdout + out.stderr


def test_real_rules_harnesses_fail_loudly_without_gpu(built, nb):
    if nb.device_count() > 0:
        pytest.skip("GPU present")
    for exe in ("nsb_selfplay_real", "nsb_usi_go_bench"):
        out = subprocess.run([os.path.join(built, exe)], capture_output=True, text=True, timeout=60)
        assert out.returncode == 2 and "no CPU fallback" in out.stderr


def test_shogi_rules_perft5_cpu(built):
    """host/rules/shogi.h against the public perft counts of the start position up to depth 5 (19,861,490 leaves) -
    the known answers that pin the move generator the search / self-play harnesses run on."""
    out = subprocess.run([os.path.join(built, "nsb_host_unit"), "--perft", "5"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_selfplay_real_rules_gpu(built):
    """Self-play with real rules and a real MCTS (host/selfplay_real.cc) on a small net: records are produced, every
    evaluated leaf has a plausible number of legal moves, no NaN rows, games progress; and with the device cache
    (raw logits stored, probabilities served: NSB_DECODE_BOTH) transpositions hit."""
    base = [os.path.join(built, "nsb_selfplay_real"), "--channels", "128", "--blocks", "2", "--batch-size", "128",
            "--frame-pool-size", "256", "--num-search-workers", "2", "--num-playouts", "16", "--seconds", "1.5", "--warmup", "0.5"]
    out = subprocess.run(base, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    assert rec["nan_rows"] == 0 and rec["evals"] > 5000 and rec["records"] > 300
    assert 15.0 <= rec["avg_legal_moves"] <= 200.0 and rec["rules"].startswith("real")
    assert 2.0 <= rec["evals"] / rec["records"] <= 20.0          # 16 playouts on full searches, 4 otherwise, + the root
    out = subprocess.run(base + ["--cache-mb", "64"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    assert rec["nan_rows"] == 0 and rec["cache_hit_rate"] > 0.02   # early-game transpositions across 256 games from hirate


@pytest.mark.gpu
def test_selfplay_writes_teacher_records_gpu(built, pkg, nb, tmp_path):
    """SaveWorker (reference src/selfplay/saveworker.cc:160-182): finished games are replayed and their full-search
    positions written as teacher records.  Tiny playout budget so that games finish within the run; the file must hold
    exactly the records the harness reports, each a shogi position with both kings, the mover's move and a winner."""
    path = str(tmp_path / "teacher.nsbt")
    out = subprocess.run([os.path.join(built, "nsb_selfplay_real"), "--channels", "128", "--blocks", "1", "--batch-size", "256",
                          "--frame-pool-size", "512", "--num-search-workers", "4", "--num-playouts", "4", "--full-search-ratio", "0.5",
                          "--seconds", "3.0", "--warmup", "0.2", "--out", path], capture_output=True, text=True, timeout=180)
    assert out.returncode == 0, out.stdout + out.stderr
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    t = rec["teacher"]
    assert t["games_saved"] > 0 and t["records_saved"] > 0, rec
    assert t["black_wins"] + t["white_wins"] + t["draws"] == t["games_saved"]
    recs = pkg.teacher_io.read_nsbt(path)
    assert len(recs) == t["records_saved"]
    assert set(np.unique(recs["winner"])) <= {0, 1, 2}
    board = recs["position"]["board"]
    assert np.all((board == 6).sum(axis=1) == 1) and np.all((board == 20).sum(axis=1) == 1)      # both kings (K = type 5)
    assert np.all(recs["position"]["side"] == recs["position"]["ply"] % 2)                        # black moves on even plies
    src = recs["from"]
    on_board = src < 81
    mover = (board[np.arange(len(recs)), np.minimum(src, 80)] - 1) // 14
    assert np.all(mover[on_board] == recs["position"]["side"][on_board])                           # the mover's own piece moves
    assert np.all(board[np.arange(len(recs)), recs["to"]][~on_board] == 0)                         # drops land on empty squares


@pytest.mark.gpu
def test_usi_go_bench_gpu(built):
    """One search tree from the start position with batched leaves under virtual loss (host/usi_go_bench.cc)."""
    out = subprocess.run([os.path.join(built, "nsb_usi_go_bench"), "--channels", "128", "--blocks", "2", "--batch-size", "128",
                          "--seconds", "1.0"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    assert rec["nodes"] > 5000 and rec["value"] > 5000 and rec["avg_batch"] > 8 and len(rec["pv"].split()) >= 2
    assert 0.0 < rec["root_win_rate"] < 1.0 and 20.0 <= rec["avg_legal_moves"] <= 120.0
