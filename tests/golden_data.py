"""Loads the reference's known answers (tests/golden/reference_known_answers.json): the single place the parity tests
take the reference's golden numbers from."""
import json
import os

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_known_answers.json")
with open(_PATH) as _f:
    GOLDEN = json.load(_f)
