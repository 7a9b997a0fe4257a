mkdir -p gpurun_out; L=gpurun_out/r2y_probe.log; : > $L
run() { echo "=== $*" >> $L; env "$@" > /tmp/p.out 2>&1; echo "rc=$?" >> /tmp/p.out; grep -v "^\[stabletts_b200\] gemm.*no error$" /tmp/p.out | tail -6 >> $L; }
run STABLETTS_B200_DEBUG=2 timeout -s KILL 70 python tests/diagnostics/probe_splitk.py 2 1316
run STABLETTS_B200_SPLITK=2 timeout -s KILL 45 python tests/diagnostics/probe_splitk.py 2 1316
run STABLETTS_B200_SPLITK=3 timeout -s KILL 45 python tests/diagnostics/probe_splitk.py 2 300
run STABLETTS_B200_SPLITK=3 timeout -s KILL 45 python tests/diagnostics/probe_splitk.py 2 1280
run STABLETTS_B200_SPLITK=1 timeout -s KILL 45 python tests/diagnostics/probe_splitk.py 1 1316
run STABLETTS_B200_SPLITK=0 timeout -s KILL 45 python tests/diagnostics/probe_splitk.py 2 1316
run STABLETTS_B200_PDL=0 timeout -s KILL 45 python tests/diagnostics/probe_splitk.py 2 1316
cat $L
