/* empty stand-in: <process.h> is a Windows header the reference's threadpool.h includes */
