#!/bin/bash
# Turns the scratch outputs of one tools/gpu_round.sh call (gpurun_out/<tag>_*) into the tracked evidence under profiles/:
# bench lines, launch list + ncu --set full summary, source-level stall lists, SASS excerpt, pytest / smoke logs, and the
# per-launch DRAM traffic that bench.py quotes (profiles/traffic.json).   usage: tools/make_profiles.sh <tag> <name>
TAG=$1; NAME=$2; G=gpurun_out; P=profiles
for f in bench_dam_break_1m.json bench_reference_arm.json launches.csv pytest.log smoke.log; do
  [ -f $G/${TAG}_$f ] && cp $G/${TAG}_$f $P/${NAME}_$f
done
ncu -i $G/${TAG}_prof.ncu-rep --page raw --csv > /tmp/${TAG}_raw.csv 2>/dev/null
{
  echo "# ${NAME} — 1M dam break K=4 (and the 4M sand pile), throughput arithmetic, tree $(git rev-parse --short HEAD)"
  echo
  echo "Commands: tools/gpu_round.sh ${TAG} (launch list: ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras; full: ncu --set full --clock-control none --import-source on -k regex:...).  The bench numbers (not under ncu) are in ${NAME}_bench_*.json."
  echo
  python tools/ncu_summary.py $G/${TAG}_launches.csv /tmp/${TAG}_raw.csv
  if [ -f $G/${TAG}_prof_sand.ncu-rep ]; then
    ncu -i $G/${TAG}_prof_sand.ncu-rep --page raw --csv > /tmp/${TAG}_sand_raw.csv 2>/dev/null
    echo; echo "## 4M sand pile"; python tools/ncu_summary.py $G/${TAG}_launches.csv /tmp/${TAG}_sand_raw.csv | sed -n '/ncu --set full/,$p'
  fi
} > $P/${NAME}_ncu_summary.md
{
  echo "# ${NAME} — where the warps stall (ncu --set full --import-source on, first captured launch, instructions with the most samples)"
  for k in k_build_table k_fluid_lambda "k_fluid_deltap<.int.1, .bool.0>"; do
    echo; echo '```'; python tools/ncu_source_stalls.py $G/${TAG}_prof.ncu-rep "$k" 60 2>/dev/null | head -45; echo '```'
  done
  if [ -f $G/${TAG}_prof_sand.ncu-rep ]; then echo; echo '```'; python tools/ncu_source_stalls.py $G/${TAG}_prof_sand.ncu-rep k_sand_iteration 80 2>/dev/null | head -45; echo '```'; fi
} > $P/${NAME}_stalls.md
{
  echo "# ${NAME} — SASS evidence (cuobjdump -sass lustrine_b200/lib/liblgpu.so): sm_100a only, bulk copies (UBLKCP) on mbarriers (SYNCS)"
  echo; echo '```'
  cuobjdump -lelf lustrine_b200/lib/liblgpu.so | head -12
  echo '```'; echo
  cuobjdump -sass lustrine_b200/lib/liblgpu.so > /tmp/${TAG}_sass.txt
  for k in 'k_fluid_lambdaILi1E' 'k_fluid_deltapILi1ELb0E' 'k_build_tableILb0ELi1E' 'k_sand_iterationI4FastLb0E' 'k_brick_desc'; do
    fn=$(grep -o "Function : _Z[0-9A-Za-z_]*${k}[0-9A-Za-z_]*" /tmp/${TAG}_sass.txt | head -1 | sed 's/Function : //')
    [ -z "$fn" ] && continue
    echo "## $(echo $fn | c++filt)"; echo
    echo "instruction mix of the whole kernel (mnemonic: count), then the bulk-copy / mbarrier lines verbatim"; echo '```'
    cuobjdump -sass -fun "$fn" lustrine_b200/lib/liblgpu.so | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sort | uniq -c | sort -rn | head -28 | awk '{printf "%s: %s   ", $2, $1} END {print ""}' | fold -w 150
    echo
    cuobjdump -sass -fun "$fn" lustrine_b200/lib/liblgpu.so | grep -E "UBLKCP|SYNCS|UTMA|FENCE.VIEW.ASYNC" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/\s*\/\*.*//' | sort | uniq -c | sort -rn | head -14
    echo '```'
  done
} > $P/${NAME}_sass.md
python - "$TAG" <<'PY'
import csv, json, sys
tag = sys.argv[1]
pipes = {}
out = {"source": "ncu --set full --clock-control none, profiles ncu summary of this round (dram__bytes_read.sum + dram__bytes_write.sum of one launch)"}
def grab(path, names):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
    res = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        for key, pat in names.items():
            if pat in name and key not in res:
                def val(m):
                    v = float(r[idx[m]].replace(",", "")); u = rows[1][idx[m]]
                    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
                res[key] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
                def pct(m):
                    try: return float(r[idx[m]].replace(",", ""))
                    except Exception: return None
                # the fp32 pipe beside the HBM figure; plus the two units that are busiest in these kernels
                pipes.setdefault(key, {"fma_pipe_pct": pct("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                                       "shared_memory_pipe_pct": pct("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                                       "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active")})
    return res
try:
    out["dam_break_1m"] = grab("/tmp/%s_raw.csv" % tag, {"k_fluid_lambda": "k_fluid_lambda", "k_fluid_deltap": "k_fluid_deltap"})
    try: out["sand_pile_4m"] = grab("/tmp/%s_sand_raw.csv" % tag, {"k_sand_iteration": "k_sand_iteration"})
    except Exception: pass
    out["pipes"] = pipes
    json.dump(out, open("profiles/traffic.json", "w"), indent=1)
    print(out)
except Exception as e:
    print("traffic.json not updated:", e)
PY
