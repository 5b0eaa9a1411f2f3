cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python - <<'PY'
import torch, time, numpy as np
import stereo_reconstruction_cv_b200 as sg
K=np.array([[1733.74,0,792.27],[0,1733.74,541.89],[0,0,1]]); P=np.array([[1700.,0,800.,0],[0,1700.,540.,0],[0,0,1,0]])
img=torch.randint(0,256,(2160,3840),dtype=torch.uint8,device='cuda')
for _ in range(3): m1,m2=sg.initUndistortRectifyMap(K,None,np.eye(3),P,(3840,2160)); o=sg.remap(img,m1,m2)
torch.cuda.synchronize()
e=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); m1,m2=sg.initUndistortRectifyMap(K,None,np.eye(3),P,(3840,2160)); e[1].record()
for _ in range(10): o=sg.remap(img,m1,m2)
e[2].record(); torch.cuda.synchronize()
print('map init ms',e[0].elapsed_time(e[1]),'remap ms',e[1].elapsed_time(e[2])/10, 'GB/s', (8.3e6*(1+8+1))/(e[1].elapsed_time(e[2])/10*1e-3)/1e9)
PY
