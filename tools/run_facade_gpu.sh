g++ -std=c++17 -O1 -Iinclude tests/cpp/test_facade.cpp -o gpurun_out/test_facade -Lpolympc_b200 -lpolympc_b200 -Wl,-rpath,/root/repo/polympc_b200
./gpurun_out/test_facade 1
./gpurun_out/test_facade 2 | tail -3
