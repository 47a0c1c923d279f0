import csv,sys,subprocess
rep,skip=sys.argv[1],sys.argv[2]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--launch-skip',skip,'--launch-count','1'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
print(rows[0][1][:120])
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)==len(hdr) and r[0].startswith('0x')]
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
S=sum(int(r[idx['# Samples']] or 0) for r in data)
print('samples',S)
top=sorted(data,key=lambda r:-int(r[idx['# Samples']] or 0))[:int(sys.argv[3]) if len(sys.argv)>3 else 25]
for r in top:
    st={c:int(r[idx[c]] or 0) for c in stall_cols}
    main=max(st,key=st.get)
    print(r[idx['# Samples']], r[idx['Instructions Executed']], r[idx['Source']].strip()[:80], main, st[main])
