import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "value %.1fG ms/step %.3f kernel_us %.1f frac %.4f e2e %.1fG p50 %.3f" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], d["e2e"]["value"]/1e9, d["e2e"]["single_query_latency_ms"]["p50"]))
