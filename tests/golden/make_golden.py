"""Generate tests/golden/mocha4.json from the reference's own fixtures.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The output is a condensed copy of chain DATA (headers, commits, validator sets, expected data
commitments) -- not reference source code -- and is committed so the tests can run anywhere.
Sources: BX/circuits/fixtures/mocha-4/*, TX/circuits/fixtures/mocha-4/*  (SURVEY 8c).
"""
import json
import os

REF = "/root/reference"
BX = f"{REF}/circuits/fixtures/mocha-4"
TX = f"{REF}/contracts/lib/tendermintx/circuits/fixtures/mocha-4"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mocha4.json")


def load(p):
    with open(p) as f:
        return json.load(f)["result"]


def main():
    g = {"source": "succinctlabs/blobstreamx@e239c6f fixtures (mocha-4)", "headers": {}, "commits": {},
         "validators": {}, "data_commitments": {}}
    # BX: headers 10000..10004 + data commitments
    for d in sorted(os.listdir(BX)):
        p = f"{BX}/{d}"
        if "-" in d:
            g["data_commitments"][d] = load(f"{p}/data_commitment.json")["data_commitment"]
        else:
            sb = load(f"{p}/signed_block.json")
            g["headers"][d] = sb["header"] if "header" in sb else sb["signed_header"]["header"]
            commit = sb.get("commit") or sb.get("signed_header", {}).get("commit")
            if commit:
                g["commits"][d] = commit
    # TX: commit.json + validators_1.json (what InputDataFetcher reads) and extra headers
    for d in sorted(os.listdir(TX)):
        p = f"{TX}/{d}"
        files = os.listdir(p)
        if "commit.json" in files:
            c = load(f"{p}/commit.json")["signed_header"]
            g["headers"][d] = c["header"]
            g["commits"][d] = c["commit"]
        elif "header.json" in files and d not in g["headers"]:
            g["headers"][d] = load(f"{p}/header.json")["header"]
        if "validators_1.json" in files:
            v = load(f"{p}/validators_1.json")
            g["validators"][d] = [
                {"address": x["address"], "pub_key": x["pub_key"]["value"], "voting_power": x["voting_power"]}
                for x in v["validators"]
            ]
    # KATs quoted in the reference's unit tests (file:line in SURVEY 8c)
    g["kats"] = {
        "tuple_height256_ff": "0" * 48 + "0000000000000100" + "ff" * 32,  # BX/circuits/builder.rs:584-605
        "tm_tree_32x48zero_root": "de8624485c0a1b8f9ecc858312916104cc3ee3ed601e405c11eaf9c5cbe05117",
        "tm_proof_depth4_root": "50d7ed02b144a75487702c9f5faaea07bb9a7385e1521e80f6080399fb9a0ffd",
        "tm_proof_depth4_aunts": [
            "78877fa898f0b4c45c9c33ae941e40617ad7c8657a307db62bc5691f92f4f60e",
            "8195d3a7e856bd9bf73464642c1e9177c7e0fbe9cf7458e2572f4e7c267676c7",
            "b1992b2f60fc8b11b83c6d9dbdd1d6abb1f5ef91c0a7aa4e7d629532048d0270",
            "0611fc80429feb4b56817f4070d289650ac0a8eaaa8975c8cc72b73e96376bff",
        ],
        "varint": [[1, [1]], [3804, [220, 29]], [1234567890, [210, 133, 216, 204, 4]],
                   [38957235239, [167, 248, 160, 144, 145, 1]],
                   [9999999999999, [255, 191, 202, 243, 132, 163, 2]],
                   [724325643436111, [207, 128, 183, 165, 211, 216, 164, 1]],
                   [9223372036854775807, [255, 255, 255, 255, 255, 255, 255, 255, 127]]],
        "validator_marshal": {
            "pubkey": "de25aec935b10f657b43fa97e5a8d4e523bdb0f9972605f0b064eff7b17048ba", "power": 100010,
            "bytes": "0a220a20de25aec935b10f657b43fa97e5a8d4e523bdb0f9972605f0b064eff7b17048ba10aa8d06"},
    }
    with open(OUT, "w") as f:
        json.dump(g, f, separators=(",", ":"), sort_keys=True)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(g["headers"]), "headers,", len(g["commits"]), "commits,",
          len(g["validators"]), "validator sets")


if __name__ == "__main__":
    main()
