for v in 1 2; do for p in fp16 fp16c; do
TB_VI_CONV2_PAIR=$v timeout 200 python bench.py --no-cpu --no-pageable --single-precision --precision $p --steps 20 > gpurun_out/pv_${v}_$p.json 2>gpurun_out/pv.err
python -c "
import json
d=json.loads(open('gpurun_out/pv_${v}_$p.json').read().strip().splitlines()[-1]); print('variant',$v,'$p',round(d['value']), d['verified'], {k:round(x['ms'],3) for k,x in d['kernels'].items() if k.startswith('conv')})"
done; done
