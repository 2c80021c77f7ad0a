"""Reads the reference's known answers for the deviatoric functions of the principal-stretch framework
(tests/src/materialresultcollection.hh: energy, dW/dlambda, and the 'second derivative' array per deformation state,
checked there to 1e-14 by tests/src/testhyperelasticity.hh:198-228) and writes them as
tests/golden/material_results.json.  Run in the build container only (/root/reference is not on the GPU box)."""
import json
import re
import sys

SRC = "/root/reference/tests/src/materialresultcollection.hh"
text = open(SRC).read()
out = {"source": "tests/src/materialresultcollection.hh (states and parameters: tests/src/testhyperelasticity.hh:28-98, 231-249)",
       "lambda": 1.37,
       "random_C": [[0.600872, -0.179083, 0.0], [-0.179083, 0.859121, 0.0], [0.0, 0.0, 1.0]],
       "functions": {}}
num = r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?"
for m in re.finditer(r"auto (\w+)Results\(\) \{(.*?)\n\}", text, re.S):
    name, body = m.group(1), m.group(2)
    states = {}
    # split into the constexpr branches
    parts = re.split(r"(?:else )?if constexpr \(def == DeformationState::(\w+)\)|\belse\b(?= \{)", body)
    # parts: [pre, state1, block1, state2, block2, ..., None, elseblock]
    i = 1
    while i < len(parts):
        state = parts[i] if parts[i] else "Undeformed"
        block = parts[i + 1]
        i += 2
        e = re.search(r"energy\s*=\s*(" + num + ")", block)
        fd = re.search(r"firstDerivatives\s*<<([^;]*);", block)
        sd = re.search(r"secondDerivatives(\.diagonal\(\))?\s*<<([^;]*);", block)
        entry = {"energy": float(e.group(1)) if e else 0.0,
                 "first": [float(v) for v in re.findall(num, fd.group(1))] if fd else [0.0, 0.0, 0.0]}
        if sd:
            vals = [float(v) for v in re.findall(num, sd.group(2))]
            if sd.group(1):
                entry["second"] = [[vals[r] if r == c else 0.0 for c in range(3)] for r in range(3)]
            else:
                assert len(vals) == 9, (name, state, vals)
                entry["second"] = [vals[0:3], vals[3:6], vals[6:9]]
        else:
            entry["second"] = [[0.0] * 3 for _ in range(3)]
        states[state] = entry
    if states:
        out["functions"][name] = states
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/material_results.json", "w"), indent=1)
print({k: sorted(v) for k, v in out["functions"].items()})
