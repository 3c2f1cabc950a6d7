"""One step of the headline circuit on specialised sweeps, for the ncu capture that profiles/r02_traffic.json is
made from, and the tool that turns the capture into that file:

    python scripts/traffic_capture.py run [n]                 # prints the kernel-set hash; run it under
        ncu --set full --clock-control none --import-source on -k regex:qj_kernel -c 11 -o gpurun_out/r02_qj_headline ...
    python scripts/traffic_capture.py record REP HASH [n]     # ncu-rep -> profiles/r02_traffic.json (here, no GPU needed)
"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n):
    import numpy as np
    from qbot_b200 import DeviceState, circuits
    gates = circuits.rc(n, 20, n)
    packed = DeviceState.pack_circuit(n, [(g.matrix(), g.target, g.controls) for g in gates])
    st = DeviceState.zero_state(n)
    st.set_jit(2)
    st.apply_circuit(packed)
    st.sync()
    st.reset_stats()
    st.apply_circuit(packed)
    st.sync()
    s = st.stats()
    print(json.dumps({"qubits": n, "kernel_set": f"{s['jit_kernel_hash']:016x}", "sweeps": s['jit_passes']}))


def record(rep, kernel_set, n):
    # REP: an .ncu-rep, or the `ncu -i REP --page raw --csv` text of one (the report itself can exceed what travels back),
    # or the --csv --log-file of a --metrics pass
    text = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(text.splitlines()) if r]
    unit = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
    tunit = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'nsecond': 1e-6}
    hdr = next(r for r in rows if r[0] == 'ID')
    how = "ncu --set full --clock-control none"
    if 'Metric Name' in hdr:
        # the log of an `ncu --metrics a,b,c --csv --log-file` pass: one row per launch and metric
        how = "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none"
        im, iu, iv = hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value')
        by = {}
        for r in rows:
            if r[0].isdigit():
                by.setdefault(int(r[0]), {})[r[im]] = (r[iu], float(r[iv].replace(',', '')))
        per, ms = [], []
        for k in sorted(by):
            (u1, v1), (u2, v2), (u3, v3) = by[k]['dram__bytes_read.sum'], by[k]['dram__bytes_write.sum'], by[k]['gpu__time_duration.sum']
            per.append(v1 * unit[u1] + v2 * unit[u2])
            ms.append(round(v3 * tunit[u3], 3))
    else:
        ir, iw, it = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
        body = rows[rows.index(hdr) + 1:]
        ur, uw = unit[body[0][ir]], unit[body[0][iw]]
        per = [float(r[ir]) * ur + float(r[iw]) * uw for r in body[1:]]
        ms = [round(float(r[it]), 3) for r in body[1:]]
    path = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    tj = json.load(open(path)) if os.path.exists(path) else {"qj_kernel": {}}
    tj['qj_kernel'][kernel_set] = {
        "qubits": n, "dram_bytes_per_launch": int(sum(per) / len(per)), "launches_captured": len(per),
        "min": int(min(per)), "max": int(max(per)), "algorithmic_bytes_per_launch": 32 << n,
        "ms_per_launch_under_ncu": ms,
        "source": (f"profiles/{os.path.basename(rep)}" if 'Metric Name' in hdr else
                   f"profiles/{os.path.basename(rep).replace('.ncu-rep', '').replace('_raw.csv', '')}_ncu_summary.txt") +
                  f" ({how}, {len(per)} launches of qj_kernel, kernel set {kernel_set})"}
    json.dump(tj, open(path, 'w'), indent=1)
    print(json.dumps(tj['qj_kernel'][kernel_set]))


if __name__ == '__main__':
    if sys.argv[1] == 'run':
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 30)
    else:
        record(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 30)
