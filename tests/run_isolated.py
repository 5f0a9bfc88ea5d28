"""Runs each selected pytest node in its own process (a CUDA fault in one kernel must not poison the rest)
and writes a summary + failure tails to gpurun_out/.  Usage: python tests/run_isolated.py <pytest args> [-j N]"""
import concurrent.futures as cf
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    args = sys.argv[1:]
    jobs = 6
    if "-j" in args:
        i = args.index("-j")
        jobs = int(args[i + 1])
        del args[i:i + 2]
    tag = "isolated"
    if "--tag" in args:
        i = args.index("--tag")
        tag = args[i + 1]
        del args[i:i + 2]
    col = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q"] + args, cwd=ROOT, capture_output=True, text=True)
    nodes = [l.strip() for l in col.stdout.splitlines() if "::" in l]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)

    def run(node):
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "--no-header", "-p", "no:cacheprovider", node],
                               cwd=ROOT, capture_output=True, text=True, timeout=900)
            return node, r.returncode, (r.stdout + r.stderr)[-3000:]
        except subprocess.TimeoutExpired:
            return node, -9, "TIMEOUT"

    results = []
    with cf.ThreadPoolExecutor(jobs) as ex:
        for node, rc, tail in ex.map(run, nodes):
            results.append({"node": node, "rc": rc, "tail": tail if rc != 0 else ""})
            print(("PASS " if rc == 0 else "FAIL ") + node, flush=True)
    with open(os.path.join(ROOT, "gpurun_out", f"{tag}.json"), "w") as f:
        json.dump(results, f, indent=1)
    bad = [r for r in results if r["rc"] != 0]
    with open(os.path.join(ROOT, "gpurun_out", f"{tag}_fail.log"), "w") as f:
        for r in bad:
            f.write("=" * 100 + "\n" + r["node"] + "\n" + r["tail"] + "\n")
    print(f"{len(results) - len(bad)} passed, {len(bad)} failed")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
