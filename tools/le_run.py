"""Development aid: local-energy timings (tools/stage_bench.le) at a few sizes; argument = number of sizes."""
import sys, json
sys.path.insert(0, "tools")
import stage_bench as sb
sizes = ((64, 512, 512), (512, 512, 512), (128, 1024, 1024), (2048, 256, 256), (100, 480, 640), (32, 2048, 2048))
for B, H, W in sizes[:int(sys.argv[1]) if len(sys.argv) > 1 else 1]:
    print(json.dumps(sb.le(B, H, W)))
