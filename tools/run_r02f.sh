mkdir -p gpurun_out
timeout 600 python tools/layout_survey.py --out gpurun_out/r02f_layout_survey.txt > /dev/null 2>&1; cat gpurun_out/r02f_layout_survey.txt
