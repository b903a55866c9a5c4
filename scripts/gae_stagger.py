import os, subprocess, sys
CHILD = open("scripts/gae_f16_timing.py").read().split("CHILD = r'''")[1].split("'''")[0]
for name, env in (("prod", {}), ("none", {"B2_GAE_TC_DEBUG": "7"}), ("none rcp->ex2", {"B2_GAE_TC_DEBUG": "7", "B2_GAE_MUFU_MODE": "1"}),
                  ("none no-mufu", {"B2_GAE_TC_DEBUG": "7", "B2_GAE_MUFU_MODE": "2"}), ("full rcp->ex2", {"B2_GAE_MUFU_MODE": "1"}),
                  ("full no-mufu", {"B2_GAE_MUFU_MODE": "2"})):
    out = subprocess.run([sys.executable, "-c", CHILD, "100000"], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
    print(name, [l for l in out.stdout.splitlines() if l.startswith("MS")] or out.stderr[-300:], flush=True)
