cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --workload infer_10s --steps 3 --warmup 2 --no-cpu-baseline --no-extra --no-e2e --profile-classes"
run() { name=$1; shift; env "$@" $B --dump-launches gpurun_out/r2_o_$name.csv > gpurun_out/r2_o_$name.json 2> gpurun_out/r2_o_$name.err; python -c "
import json,csv
d=json.load(open('gpurun_out/r2_o_$name.json'))
rows=list(csv.DictReader(open('gpurun_out/r2_o_$name.csv')))
tags=[r['tag'] for r in rows]
idx=max(i for i,t in enumerate(tags) if t=='conv_pre:fwd')
sel={r['tag']:float(r['ms'])*1000 for r in rows[idx:] if r['tag']}
pick=['resblocks.6.convs1.0:fwd','resblocks.6.convs2.0:fwd','resblocks.6.convs2.2:fwd','resblocks.8.convs1.0:fwd','resblocks.9.convs1.0:fwd','resblocks.9.convs2.0:fwd','resblocks.9.convs2.2:fwd','resblocks.11.convs1.0:fwd','resblocks.3.convs2.2:fwd','resblocks.5.convs1.0:fwd']
print('$name', round(d['ms_per_step'],2), [round(sel[p]) for p in pick])"; }
run base A=1
run na8 VCD_CONV_NA_SMALL=8
run na12 VCD_CONV_NA_SMALL=12 VCD_CONV_NE=4
run occ2 VCD_CONV_OCC2=1
run occ2_smem100 VCD_CONV_OCC2=1 VCD_CONV_SMEM_KB=100
run mt2 VCD_CONV_MT_SMALL=2
run mt4 VCD_CONV_MT_SMALL=4
run ne4 VCD_CONV_NE=4
run nolean VCD_CONV_LEAN=0
run noesmem VCD_CONV_ESMEM=0
