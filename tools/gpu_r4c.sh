#!/bin/bash
# A/B of the row pipelines: SSW_ROW_PIPE = 0 (RowFwd / RowInv) / 1 on c2 and c3; bit-identity check of the two first
TAG=${1:-r4c}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python - > $OUT/rowpipe_check_$TAG.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
res = {}
for v in ('0', '1'):
    os.environ['SSW_ROW_PIPE'] = v
    import spread_spectrum_watermarking_b200 as wm
    from spread_spectrum_watermarking_b200._lib import lib, check
    ctx = wm.Context(0)
    for (w, h, b) in ((3840, 2160, 2), (1920, 1080, 5), (640, 444, 3), (1280, 720, 2), (2560, 1440, 1), (1080, 1920, 2), (2160, 3840, 1)):
        fr = torch.empty((b, h, w, 3), dtype=torch.uint8, device='cuda')
        check(lib.ssw_synth_frame_rgb8_dev(ctx.handle, w, h, 2, 0, b, fr.data_ptr()))
        pl = torch.empty((b, h, w), dtype=torch.float32, device='cuda')
        torch.cuda.synchronize()
        try:
            check(lib.ssw_stage_forward_rgb8_dev(ctx.handle, fr.data_ptr(), w, h, b, pl.data_ptr()))
            ctx.synchronize()
            out = torch.empty_like(fr)
            pl2 = pl.clone()
            torch.cuda.synchronize()
            check(lib.ssw_stage_inverse_rgb8_dev(ctx.handle, pl2.data_ptr(), fr.data_ptr(), w, h, b, out.data_ptr()))
            ctx.synchronize()
            res[(v, w, h)] = (pl.cpu().numpy(), out.cpu().numpy(), fr.cpu().numpy())
            print('variant', v, (w, h, b), 'ok, |c|max %.4g, roundtrip max |d| %d' % (float(pl.abs().max()), int((out.int() - fr.int()).abs().max())), flush=True)
        except Exception as e:
            print('variant', v, (w, h, b), 'FAILED', e, flush=True)
    ctx.close()
for (v, w, h), (pl, out, fr) in res.items():
    if v == '0': continue
    base = res.get(('0', w, h))
    if base is None: continue
    print('variant', v, (w, h), 'planes identical:', bool(np.array_equal(pl, base[0])), ' rgb8 identical:', bool(np.array_equal(out, base[1])),
          ' max rel diff %.3g' % (np.abs(pl - base[0]).max() / np.abs(base[0]).max()))
PY
cat $OUT/rowpipe_check_$TAG.log
for rep in 1 2; do
for v in 0 1; do
  SSW_ROW_PIPE=$v timeout 300 python bench.py --workload c2 --no-cpu-baseline --no-e2e > $OUT/bench_c2_${TAG}_p${v}_$rep.json 2> $OUT/bench_c2_${TAG}_p${v}_$rep.err; echo "c2 rowpipe=$v rc=$?"; tail -2 $OUT/bench_c2_${TAG}_p${v}_$rep.err
done
done
for v in 0 1; do
  SSW_ROW_PIPE=$v timeout 300 python bench.py --workload c3 --no-cpu-baseline --no-e2e --steps 20 > $OUT/bench_c3_${TAG}_p$v.json 2> $OUT/bench_c3_${TAG}_p$v.err; echo "c3 rowpipe=$v rc=$?"; tail -2 $OUT/bench_c3_${TAG}_p$v.err
done
timeout 300 python bench.py --workload c5 --no-cpu-baseline --no-e2e > $OUT/bench_c5_${TAG}.json 2> $OUT/bench_c5_${TAG}.err; echo "c5 rc=$?"; tail -2 $OUT/bench_c5_${TAG}.err
python tools/kernels_table.py $OUT/bench_c2_${TAG}_p*.json $OUT/bench_c3_${TAG}_p*.json $OUT/bench_c5_${TAG}.json 2>&1 | grep -E "json|fwd_|inv_|similarity_bank|topk"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu_$TAG.log
