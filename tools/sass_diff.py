"""Per-function SASS comparison of two object files (refactors that must not change device code):

    python tools/sass_diff.py old.o new.o [--map 's/OLD_REGEX/NEW/']...

Every function of old.o is looked up in new.o (after the optional name rewrites, e.g. for an added template argument)
and its instruction stream (addresses stripped) compared.  Exit status 1 if any function differs or is missing.
Used in round 1 for the CTA-size template of radix_onesweep_kernel and the removal of the rejected scan variants."""
import hashlib, re, subprocess, sys


def funcs(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    res, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1); res[name] = hashlib.md5(); continue
        if name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            res[name].update(re.sub(r"/\*[0-9a-f]*\*/", "", line).encode())
    return {k: v.hexdigest() for k, v in res.items()}


def main():
    args = sys.argv[1:]
    maps = []
    while "--map" in args:
        i = args.index("--map"); spec = args[i + 1]; del args[i:i + 2]
        _, a, b_, *_ = spec.split(spec[1]) if spec.startswith("s") else (None, *spec.split("=>"))
        maps.append((re.compile(a), b_))
    old, new = funcs(args[0]), funcs(args[1])
    same = diff = missing = 0
    for name, h in old.items():
        n2 = name
        for a, b_ in maps:
            n2 = a.sub(b_, n2)
        if n2 not in new:
            missing += 1; print("MISSING", name)
        elif new[n2] == h:
            same += 1
        else:
            diff += 1; print("DIFF   ", name)
    print(f"same {same}  diff {diff}  missing {missing}  only-in-new {len(new) - same - diff}")
    sys.exit(1 if diff or missing else 0)


if __name__ == "__main__":
    main()
