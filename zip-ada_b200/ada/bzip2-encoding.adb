--  BZip2.Encoding - replacement BODY that routes Encode to the b2gpu CUDA library.
--
--  The SPEC (zip_lib/bzip2-encoding.ads of Zip-Ada v.62) is kept byte for byte: same generic
--  formals (Read_Byte, More_Bytes, Write_Byte), same Encode (option, size_hint).  This body only
--  drains the input callbacks into a buffer, calls the C ABI of include/b2gpu.h, and replays the
--  result through Write_Byte in order.  Exceptions raised inside the callbacks (User_abort,
--  Compression_inefficient, ...) propagate through Encode; the buffers are released and the
--  handle goes back to the pool in an exception handler, mirroring the clean-up of the original
--  body (bzip2-encoding.adb:1128-1133, :1351-1357, :1377-1381).
--
--  Handles (CUDA streams, tables, device workspaces) are expensive to create, and Zip.Create calls
--  Encode once per archive entry (zip-create.adb:253-265): they are kept in a protected pool, one
--  per (block size, concurrent caller).  Several tasks may run Encode at the same time
--  (doc/zipada.txt:26); each takes its own handle from the pool.
--
--  NOTE: there is no Ada compiler in the build image of this repository, so this file has not
--  been compiled.  It is deliberately small; the C++ mirror host/bzip2_encoding.hpp has the same
--  logic (pool included) and is what the tests drive.  Link with -lb2gpu.

with Ada.Unchecked_Deallocation;
with Interfaces.C;
with System;

package body BZip2.Encoding is

  use Interfaces, Interfaces.C;
  use type System.Address;

  subtype Handle is System.Address;

  function b2_create (level, device : int; enc : access Handle) return int;
  pragma Import (C, b2_create, "b2_create");

  procedure b2_destroy (enc : Handle);
  pragma Import (C, b2_destroy, "b2_destroy");
  pragma Unreferenced (b2_destroy);  --  pooled handles live as long as the program

  function b2_bound (n : Unsigned_64) return Unsigned_64;
  pragma Import (C, b2_bound, "b2_bound");

  function b2_encode_stream
    (enc       : Handle;
     input     : System.Address;
     n         : Unsigned_64;
     size_hint : Integer_64;
     output    : System.Address;
     out_cap   : Unsigned_64;
     out_len   : access Unsigned_64) return int;
  pragma Import (C, b2_encode_stream, "b2_encode_stream");

  type Byte_Array is array (Unsigned_64 range <>) of aliased Byte;
  type Byte_Array_Access is access Byte_Array;
  procedure Free is new Ada.Unchecked_Deallocation (Byte_Array, Byte_Array_Access);

  b2gpu_error : exception;

  Device : constant int := 0;  --  CUDA device of the pooled handles

  --  Idle handles, per block size.  Take returns Null_Address when none is idle.
  type Handle_List is array (1 .. 64) of Handle;
  type Pool_Row is record
    idle  : Handle_List := (others => System.Null_Address);
    count : Natural := 0;
  end record;
  type Pool_Table is array (Compression_Option) of Pool_Row;

  protected Pool is
    procedure Take (option : Compression_Option; h : out Handle);
    procedure Give (option : Compression_Option; h : Handle; kept : out Boolean);
  private
    rows : Pool_Table;
  end Pool;

  protected body Pool is
    procedure Take (option : Compression_Option; h : out Handle) is
    begin
      if rows (option).count = 0 then
        h := System.Null_Address;
      else
        h := rows (option).idle (rows (option).count);
        rows (option).count := rows (option).count - 1;
      end if;
    end Take;
    procedure Give (option : Compression_Option; h : Handle; kept : out Boolean) is
    begin
      kept := rows (option).count < rows (option).idle'Last;
      if kept then
        rows (option).count := rows (option).count + 1;
        rows (option).idle (rows (option).count) := h;
      end if;
    end Give;
  end Pool;

  procedure Encode
    (option    : Compression_Option := block_900k;
     size_hint : Stream_Size_Type   := unknown_size)
  is
    level : constant int :=
      (case option is
         when block_100k => 1,
         when block_400k => 4,
         when block_900k => 9);
    enc     : aliased Handle := System.Null_Address;
    inp     : Byte_Array_Access := new Byte_Array (1 .. 1_048_576);
    outp    : Byte_Array_Access := null;
    n       : Unsigned_64 := 0;
    out_len : aliased Unsigned_64 := 0;
    grown   : Byte_Array_Access;
    kept    : Boolean;
    procedure b2_destroy_now (e : Handle);
    pragma Import (C, b2_destroy_now, "b2_destroy");
  begin
    --  1) Drain the input callbacks (they update the Zip CRC-32 and the feedback, and may raise
    --     User_abort, zip-compress-bzip2_e.adb:70-106, exactly as before).
    while More_Bytes loop
      if n = inp'Last then
        grown := new Byte_Array (1 .. inp'Last * 2);
        grown (1 .. n) := inp (1 .. n);
        Free (inp);
        inp := grown;
      end if;
      n := n + 1;
      inp (n) := Read_Byte;
    end loop;
    --  2) One call to the device encoder (whole stream: header, blocks, footer) on a pooled handle.
    Pool.Take (option, enc);
    if enc = System.Null_Address and then b2_create (level, Device, enc'Access) /= 0 then
      raise b2gpu_error with "b2_create failed (no CUDA device?) - there is no CPU fallback";
    end if;
    outp := new Byte_Array (1 .. b2_bound (n) + 1024 * (n / 40_000 + 16));
    if b2_encode_stream
         (enc, inp (1)'Address, n, Integer_64 (size_hint),
          outp (1)'Address, outp'Length, out_len'Access) /= 0
    then
      raise b2gpu_error with "b2_encode_stream failed";
    end if;
    Pool.Give (option, enc, kept);
    if not kept then
      b2_destroy_now (enc);
    end if;
    enc := System.Null_Address;
    Free (inp);
    --  3) Replay the result through Write_Byte, in order (may raise Compression_inefficient,
    --     zip-compress.adb:480-486: the handle is already back in the pool).
    for i in 1 .. out_len loop
      Write_Byte (outp (i));
    end loop;
    Free (outp);
  exception
    when others =>
      if enc /= System.Null_Address then
        Pool.Give (option, enc, kept);
        if not kept then
          b2_destroy_now (enc);
        end if;
      end if;
      Free (inp);
      Free (outp);
      raise;
  end Encode;

end BZip2.Encoding;
