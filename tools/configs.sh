# BASELINE.json configs 1-4 as diagnostic bench lines (config 5 is the default bench.py line)
echo "C1 SU(2) 16^4 fp32:";       python bench.py --group 2 --precision 0 --dims 16x16x16x16 --steps 50 --warmup 10 --no-cpu | cut -c1-260
echo "C2 SU(3) 24^4 fp64 + 4 OR:"; python bench.py --dims 24x24x24x24 --or-steps 4 --steps 20 --warmup 5 --no-cpu | cut -c1-330
echo "C2 SU(3) 24^4 fp64 HB only:"; python bench.py --dims 24x24x24x24 --steps 20 --warmup 5 --no-cpu | cut -c1-260
echo "C3 SU(3) 32^3x8 mixed, PHI=3:"; python bench.py --precision 2 --dims 32x32x32x8 --phi 3.0 --steps 30 --warmup 5 --no-cpu | cut -c1-260
echo "C4 O(4) 32^4:";             python bench.py --on --steps 20 --warmup 5 --no-cpu | cut -c1-260
echo "C5 hot start:";             python bench.py --hot --steps 20 --warmup 10 --no-cpu | cut -c1-260
echo "64^4 fp32:";                python bench.py --precision 0 --steps 20 --warmup 5 --no-cpu | cut -c1-260
echo "64^4 SU(2) fp64:";          python bench.py --group 2 --steps 20 --warmup 5 --no-cpu | cut -c1-260
echo "64^4 + 4 OR:";              python bench.py --or-steps 4 --steps 5 --warmup 2 --no-cpu | cut -c1-330
