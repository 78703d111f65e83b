"""Regenerates tests/golden/notebook_kats.json from the reference notebooks.

Runs only in the build container (reads /root/reference, which does not exist
on the GPU box).  The JSON holds numbers PRINTED by the reference's own
notebooks - the only known answers the reference has for this path.
The K-A series / evaluation times are parsed from the notebook outputs; the
scalar KATs and the printed state entries were transcribed from the cells named
in each entry's "source" field.
"""
import json
import re

NB = "/root/reference/docs/basic_usage.ipynb"


def main() -> None:
    nb = json.load(open(NB))
    series = "".join(nb["cells"][25]["outputs"][0]["data"]["text/plain"])
    vals = [float(x) for x in re.findall(r"-?\d+\.\d+", series.split("dtype")[0])]
    times = "".join(nb["cells"][14]["outputs"][0]["text"]).split("Wavefunctions")[0]
    tv = [float(x) for x in re.findall(r"\d+\.\d+", times.split("dtype")[0])]
    gold = json.load(open("tests/golden/notebook_kats.json"))
    gold["K-A"]["sum_z"], gold["K-A"]["eval_times"] = vals, tv
    json.dump(gold, open("tests/golden/notebook_kats.json", "w"), indent=1)


if __name__ == "__main__":
    main()
