"""summarise_ncu.py REPORT.ncu-rep WARP_VISITS [MARKER ...] -- text summary of one captured kernel: key raw metrics, stall reasons per issue, and the
warp-state samples of the ncu source page cut into code regions at the given SASS markers (default: the barrier / tensor-memory / SFU landmarks of
k_sweep_tma_psi).  Run where ncu is installed (no GPU needed): python profiles/summarise_ncu.py gpurun_out/x.ncu-rep 704512 > profiles/r02_x.txt"""
import csv, subprocess, sys

rep, warp_visits = sys.argv[1], float(sys.argv[2])
markers = sys.argv[3:] or ["BAR.SYNC", "SYNCS.PHASECHK", "UBLKCP", "STTM", "UTCBAR", "LDTM", "MUFU.RSQ", "MUFU.EX2", "STG.E.128", "EXIT"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
print("kernel:", m.get("Kernel Name", ("?", ""))[0])
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
for k in keys:
    if k in m:
        print(f"  {k:95s} {m[k][0]} {m[k][1]}")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and float(m[h][0]) > 0.02:
        print(f"  {h:95s} {float(m[h][0]):.3f}")
inst = float(m["smsp__inst_executed.sum"][0])
dr = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
dram = sum(float(m[k][0]) * dr[m[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
tu = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
t = float(m["gpu__time_duration.sum"][0]) * tu[m["gpu__time_duration.sum"][1]]
print(f"\nderived: {inst / warp_visits:.0f} warp-instructions per warp visit; DRAM {dram / (32 * warp_visits):.1f} B per site visit (algorithmic 64 B: state in + state out) at "
      f"{dram / t / 1e12:.2f} TB/s; {t / (32 * warp_visits) * 1e12:.1f} ps per site visit under ncu (cold caches, serialised)")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print(f"\nwarp-state samples by code region (ncu source page; a region ends at the named instruction; {tot} samples):")
print(f"{'from':>6} {'to':>6} {'instr':>5} {'exec / warp visit':>17} {'samples':>8}  top stall reasons | last instruction")
acc = n = ex = 0
agg = {h: 0 for h in stalls}
start = data[0][ix["Address"]][-4:]
for r in data + [None]:
    if r is not None:
        acc += int(r[ix["# Samples"]] or 0); n += 1; ex += int(r[ix["Instructions Executed"]] or 0)
        for h in stalls:
            agg[h] += int(r[ix[h]] or 0)
    if r is None or any(k in r[ix["Source"]] for k in markers):
        top = " ".join(f"{h[6:]}:{v}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:4] if v)
        end = r[ix["Address"]][-4:] if r is not None else "end"
        print(f"{start:>6} {end:>6} {n:5d} {ex / warp_visits:17.1f} {100 * acc / tot:7.2f}%  {top} | {(r[ix['Source']][:48] if r is not None else '')}")
        acc = n = ex = 0; agg = {h: 0 for h in stalls}; start = end
print("\nhottest instructions:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:10]:
    s = int(r[ix["# Samples"]]); st = {h: int(r[ix[h]] or 0) for h in stalls}
    print(f"  {r[ix['Address']][-4:]} {100 * s / tot:5.2f}%  {r[ix['Source']][:64]:64s} {sorted(st.items(), key=lambda kv: -kv[1])[0]}")
