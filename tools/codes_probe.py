import sys, json
sys.path.insert(0, "/root/repo/tools")
import sweep
print(json.dumps(sweep.codes_case(4096, "flint", False)))
print(json.dumps(sweep.codes_case(4096, "flint", True)))
