#!/usr/bin/env python
"""Turn the captures of tools/refresh_profiles.sh (gpurun_out/final/) into the files under profiles/ (build container).

    python tools/make_profiles.py [gpurun_out/final]

The by-line tools need the cubins of THE SAME build that was profiled: they are extracted here from
light_transport_b200/liblt_b200.so, so run this before rebuilding the library."""
import csv, glob, json, os, shutil, subprocess, sys, tempfile
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SRC = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "final")
PROF = os.path.join(ROOT, "profiles")
py = sys.executable


def run(*a):
    return subprocess.run(a, capture_output=True, text=True).stdout


def metric_rows(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    out = {}
    for r in rows[1:]:
        out.setdefault((r[ix["ID"]], r[ix["Kernel Name"]]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    return out


def rays_per_launch(bench_json):
    d = json.load(open(bench_json))
    return d["value"] * 1e6 * d["ms_per_step"] * 1e-3


def main():
    cub = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "light_transport_b200", "liblt_b200.so")], cwd=cub, capture_output=True)
    f32 = glob.glob(os.path.join(cub, "lt_walk_f32*.cubin"))[0]
    f64 = glob.glob(os.path.join(cub, "lt_walk_f64*.cubin"))[0]
    caps = [("c2walk32", "r02_walk_compat_f32", f32, "walk_megakernelIfLi0ELi0ELi1E",
             "kernel lt::walk_megakernel<float, COMPAT, REFERENCE, PH=1> on the LTS scene, probe 512x512 x 64 spp, D=16\n"
             "command: python tools/run_once.py megakernel compat f32 512 64   (second launch)"),
            ("c2walk64", "r02_walk_compat_f64", f64, "walk_megakernelIdLi0ELi0ELi1E",
             "kernel lt::walk_megakernel<double, COMPAT, REFERENCE, PH=1> (the like-for-like float64 arm) on the LTS scene, probe 512x512 x 64 spp, D=16\n"
             "command: python tools/run_once.py megakernel compat f64 512 64   (second launch)"),
            ("c3lbvh2", "r02_c3_lbvh2", f32, "walk_megakernelIfLi3ELi1ELi1E",
             "kernel lt::walk_megakernel<float, lbvh2, ITERATIVE> on the C3 scene (1,000,028 triangles), 1920x1080 x 4 spp, D=8\n"
             "command: python tools/run_mesh.py --spp 4 --reps 2 --trav lbvh2   (second launch)")]
    for cap, name, cubin, kern, head in caps:
        rep = os.path.join(SRC, cap + ".ncu-rep")
        if not os.path.exists(rep):
            print("missing", rep); continue
        head = "ncu --set full --clock-control none --import-source on, " + head + "\n\n"
        open(os.path.join(PROF, name + "_ncu_summary.txt"), "w").write(head + run(py, os.path.join(ROOT, "tools", "ncu_summary.py"), rep))
        open(os.path.join(PROF, name + "_by_source_line.txt"), "w").write(run(py, os.path.join(ROOT, "tools", "ncu_by_line.py"), rep, cubin, kern, "60"))
        open(os.path.join(PROF, name + "_stalls.txt"), "w").write(run(py, os.path.join(ROOT, "tools", "ncu_stalls.py"), rep, cubin, kern))
        print("wrote", name)
    # ---- traffic
    tpath = os.path.join(PROF, "r02_traffic.json")
    tj = json.load(open(tpath))
    c2csv, c3csv = os.path.join(SRC, "r02_launches_bench_c2.csv"), os.path.join(SRC, "r02_traffic_bench_c3.csv")
    if os.path.exists(c2csv):
        shutil.copy(c2csv, os.path.join(PROF, "r02_launches_bench_c2.csv"))
        walks = [m for (i, k), m in metric_rows(c2csv).items() if "walk_megakernel" in k]
        big = max(walks, key=lambda m: m["gpu__time_duration.sum"])
        tj["c2:compat:megakernel:f32"].update(bytes=int(big["dram__bytes_read.sum"] + big["dram__bytes_write.sum"]), read=int(big["dram__bytes_read.sum"]),
                                              write=int(big["dram__bytes_write.sum"]), kernel_ns_under_ncu=int(big["gpu__time_duration.sum"]))
    if os.path.exists(c3csv):
        shutil.copy(c3csv, os.path.join(PROF, "r02_traffic_bench_c3.csv"))
        walks = [m for (i, k), m in metric_rows(c3csv).items() if "walk_megakernel" in k]
        big = max(walks, key=lambda m: m["gpu__time_duration.sum"])
        rays = rays_per_launch(os.path.join(SRC, "plain_c3.json"))
        sect = big["lts__t_sectors_srcunit_tex_op_read.sum"]
        tj["c3:lbvh2:megakernel:f32"].update(bytes=int(big["dram__bytes_read.sum"] + big["dram__bytes_write.sum"]), read=int(big["dram__bytes_read.sum"]),
                                             write=int(big["dram__bytes_write.sum"]), kernel_ns_under_ncu=int(big["gpu__time_duration.sum"]),
                                             l2_read_sectors=int(sect), l2_read_bytes_per_ray=sect * 32 / rays,
                                             dram_bytes_per_ray=(big["dram__bytes_read.sum"] + big["dram__bytes_write.sum"]) / rays, rays_per_launch=int(rays))
    json.dump(tj, open(tpath, "w"), indent=1)
    print(json.dumps({k: v for k, v in tj.items() if k != "note"}, indent=1)[:1500])


if __name__ == "__main__":
    main()
