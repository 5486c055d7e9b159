#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference checkout (run in the build container only).

Inputs  : /root/reference/tests/fastq/*            (the reference's own FASTQ fixtures, 53 B - 1.4 KB)
          /root/reference/docs/fq-count.md:27-43   (the only golden `sc fq-count` rows the reference holds)
          /root/reference/docs/fq-meta.md:32-37    (the only golden min_qual/max_qual/n_lines rows)
Outputs : tests/golden/fastq/*                     byte-identical copies of the fixtures (test DATA, not source)
          tests/golden/fq_count_docs.tsv           file, reads, gc_content (as printed in the docs, 6 sig. digits),
                                                   gc_bases, n_bases, bases
          tests/golden/fq_meta_docs.tsv            file, qual_format, qual_phred, qual_multiple, min_qual, max_qual, n_lines
The GPU box has no /root/reference, so tests read only the committed outputs.
"""
import os
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def md_rows(path):
    rows = []
    for line in open(path, encoding="utf-8"):
        line = line.strip()
        if not line.startswith("|"):
            continue
        cells = [c.strip() for c in line.strip("|").split("|")]
        if all(set(c) <= set("-: ") for c in cells):
            continue
        rows.append(cells)
    return rows


def main():
    src = os.path.join(REF, "tests", "fastq")
    dst = os.path.join(HERE, "fastq")
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(src)):
        shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))

    rows = md_rows(os.path.join(REF, "docs", "fq-count.md"))
    head, body = rows[0], rows[1:]
    idx = {h: i for i, h in enumerate(head)}
    with open(os.path.join(HERE, "fq_count_docs.tsv"), "w") as f:
        f.write("file\treads\tgc_content\tgc_bases\tn_bases\tbases\n")
        for r in body:
            f.write("\t".join([r[idx["basename"]], r[idx["reads"]], r[idx["gc_content"]], r[idx["gc_bases"]],
                               r[idx["n_bases"]], r[idx["bases"]]]) + "\n")

    rows = md_rows(os.path.join(REF, "docs", "fq-meta.md"))
    head, body = rows[0], rows[1:]
    idx = {h: i for i, h in enumerate(head)}
    with open(os.path.join(HERE, "fq_meta_docs.tsv"), "w") as f:
        f.write("file\tqual_format\tqual_phred\tqual_multiple\tmin_qual\tmax_qual\tn_lines\n")
        for r in body:
            f.write("\t".join([r[idx["basename"]], r[idx["qual_format"]], r[idx["qual_phred"]], r[idx["qual_multiple"]],
                               r[idx["min_qual"]], r[idx["max_qual"]], r[idx["n_lines"]]]) + "\n")


if __name__ == "__main__":
    main()
