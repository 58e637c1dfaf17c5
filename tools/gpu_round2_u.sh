set -x
timeout 1500 python tools/certify_rows.py gpurun_out/r2_certify_rows.md 1 > gpurun_out/r2u_certify_rows.log 2>&1
tail -25 gpurun_out/r2u_certify_rows.log
