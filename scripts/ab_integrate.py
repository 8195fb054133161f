"""A/B helper: prints the integrate-kernel time for the library selected by OPB_LIB_PATH."""
import json, os, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--steps", "100", "--warmup", "5", "--no-cpu-baseline"], capture_output=True, text=True)
try:
    j = json.loads(out.stdout.strip().splitlines()[-1])
    r = j["roofline"]
    print(os.environ.get("OPB_LIB_PATH", "default"), "value %.0f fps  k2 %.1f us  sel %.1f us  frac %.3f  e2e %.0f" % (j["value"], r["kernel_ms"] * 1e3, r["select_ms"] * 1e3, r["frac"], j["e2e"]["value"]))
except Exception as e:
    print("failed", e, out.stdout[-500:], out.stderr[-2000:])
