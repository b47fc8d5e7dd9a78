import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d.get("roofline", {})
    print(f"{f.split('/')[-1]:22s} {d['value']:9.0f} Mrays/s {d['ms_per_step']:9.2f} ms/step e2e {d['e2e']['value']:8.0f} N={d['n_gpus']} "
          f"frac {r.get('frac', 0):.3f} rays/path {d.get('rays_per_path', 0):.2f} box {r.get('box_tests_per_ray', 0):.1f} tri {r.get('tri_tests_per_ray', 0):.2f} "
          f"util {r.get('lane_utilisation', 0):.3f} cpu {(d.get('cpu_baseline') or {}).get('value')} rmse {d.get('time_to_rmse')}")
