import re, sys, subprocess
for log in sys.argv[1:]:
    txt = open(log).read()
    ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt)
    names = [e[0] for e in ents]
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for d, e in zip(dem, ents):
        d = re.sub(r"ce::\(anonymous namespace\)::", "", d)
        d = re.sub(r"\(.*", "", d)
        print("%-90s regs=%3s stack=%4s spill_st=%4s spill_ld=%4s" % (d[:90], e[4], e[1], e[2], e[3]))
